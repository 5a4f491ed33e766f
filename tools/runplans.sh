W=${1:-C4}
for plan in "" "33,1,1"; do
  PVT_PLAN="$plan" PVT_DEBUG_PLAN=1 timeout 300 python bench.py --workload $W --steps 24 --warmup 4 --no-cpu > /tmp/b.json 2> /tmp/b.err
  echo "plan=[$plan] rc=$? $(grep -m1 'plan:' /tmp/b.err)"
  python -c "
import json
d=json.load(open('/tmp/b.json'))
print('   ms/step %.4f  ncc %.4f ms  frac %.3f' % (d['ms_per_step'], d['kernel_ms_per_step']['ncc'], d['roofline']['frac']))"
done
