set -x; mkdir -p gpurun_out
python -m pytest tests/test_winstats.py -x -q > gpurun_out/t_winstats.log 2>&1; tail -3 gpurun_out/t_winstats.log
python -m pytest tests -m gpu -x -q > gpurun_out/gputest4.log 2>&1; tail -3 gpurun_out/gputest4.log
python bench.py --no-cpu --steps 200 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -c 400 gpurun_out/bench4.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_winstats -s 6 -c 1 -f -o gpurun_out/winstats_c4 python bench.py --workload C4 --steps 4 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_ws.log 2>&1; tail -2 gpurun_out/ncu_ws.log
