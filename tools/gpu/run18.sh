set -x; mkdir -p gpurun_out
timeout 120 python tools/timeline.py C2
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/gputest18.log 2>&1; tail -5 gpurun_out/gputest18.log
timeout 600 python bench.py --no-cpu --steps 600 > gpurun_out/bench18.json 2> gpurun_out/bench18.err; tail -c 300 gpurun_out/bench18.err
