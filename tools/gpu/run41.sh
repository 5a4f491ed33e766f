set -x; mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "gray or ingest or G6 or strided or to_gray" 2>&1 | tail -3
python - <<'PY'
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--no-cpu", "--no-extra", "--steps", "40", "--warmup", "4"], capture_output=True, text=True).stdout
d = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
print("ingest", d["ingest"]["achieved"], d["ingest"]["frac"], d["ingest"]["us_per_launch"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_ingest$ -s 10 -c 1 -f -o gpurun_out/ingest_r2 python bench.py --steps 40 --warmup 4 --no-cpu --no-extra > gpurun_out/ncu_ingest.log 2>&1; tail -1 gpurun_out/ncu_ingest.log
