set -x; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_step.py -x -q > gpurun_out/t_fused.log 2>&1; tail -15 gpurun_out/t_fused.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest8.log 2>&1; tail -3 gpurun_out/gputest8.log
timeout 600 python bench.py --no-cpu --steps 600 > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -c 400 gpurun_out/bench8.err
timeout 120 python tools/timeline.py C2
