set -x; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/gputest21.log 2>&1; tail -4 gpurun_out/gputest21.log
timeout 900 python bench.py > gpurun_out/bench21.json 2> gpurun_out/bench21.err; tail -c 300 gpurun_out/bench21.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench21_ref.json 2> gpurun_out/bench21_ref.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ncc_local -s 40 -c 1 -f -o gpurun_out/local_c2 python bench.py --workload C2 --steps 40 --warmup 8 --no-cpu --no-extra > gpurun_out/ncu_local.log 2>&1; tail -1 gpurun_out/ncu_local.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ncc_tc -s 6 -c 1 -f -o gpurun_out/tc_c5_after python bench.py --workload C5 --kernel tc --steps 4 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_tc2.log 2>&1; tail -1 gpurun_out/ncu_tc2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_default_r2.csv python bench.py --steps 40 --warmup 4 --no-cpu --no-extra > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
