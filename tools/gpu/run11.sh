set -x; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_local_search.py -x -q > gpurun_out/t_local.log 2>&1; tail -15 gpurun_out/t_local.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gputest11.log 2>&1; tail -8 gpurun_out/gputest11.log
timeout 120 python tools/timeline.py C2
timeout 600 python bench.py --no-cpu --steps 600 > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -c 400 gpurun_out/bench11.err
