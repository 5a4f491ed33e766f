set -x; mkdir -p gpurun_out
python - <<'PY'
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--no-cpu", "--no-extra", "--steps", "40", "--warmup", "4"], capture_output=True, text=True)
d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
print("ingest", json.dumps({k: d["ingest"][k] for k in ("achieved", "frac", "us_per_launch", "us_per_launch_event_nodes", "frac_event_nodes")}))
PY
