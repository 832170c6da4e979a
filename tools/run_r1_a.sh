set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r1_gputests_b.log
for k in rbf linear; do timeout 600 python bench.py --kernel $k --steps 5 --warmup 3 > gpurun_out/r1_bench_cfg4_$k.json 2> gpurun_out/r1_bench_cfg4_$k.err; done
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_cfg2.json 2> gpurun_out/r1_bench_cfg2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_cfg4.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sigkern_fo_tma -s 20 -c 2 -f -o gpurun_out/r1_prof_recursion python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:delta_producer -s 20 -c 2 -f -o gpurun_out/r1_prof_producer python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r1_ncu_full_prod.log 2>&1
tail -3 gpurun_out/r1_gputests_b.log; cat gpurun_out/r1_bench_cfg4_rbf.json gpurun_out/r1_bench_cfg4_linear.json gpurun_out/r1_bench_cfg2.json; tail -3 gpurun_out/*.err
