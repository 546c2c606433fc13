# round 2: one GPU call with everything a single GPU can give (each step with its own timeout, outputs in gpurun_out/)
O=gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -45 > $O/r2_tests_full.log
timeout -k 10 300 python bench.py --steps 5 --warmup 3 > $O/r2_bench_a.json 2> $O/r2_bench_a.err
WPT_TRACE_PHOTONS=1 timeout -k 10 100 python scripts/run_configs.py 3 3m > $O/r2_cfg3.log 2>&1
timeout -k 10 300 bash scripts/r2_gpu_b.sh > $O/r2_perf_b.log 2>&1
# ncu: launch list of the bench command, then one full capture of the headline kernel and of the museum kernel
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-target > $O/r2_ncu_launch.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_mega -s 3 -c 1 -o $O/r2_mega_bench -f python bench.py --steps 1 --warmup 3 --no-cpu --no-target > $O/r2_ncu_full.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_mega -s 1 -c 1 -o $O/r2_mega_museum -f python scripts/time_step.py 8 1 2 1 0 0 > $O/r2_ncu_museum.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_mega -s 1 -c 1 -o $O/r2_mega_bvh4_pnee -f python scripts/time_step.py 16 1 4 2 0 > $O/r2_ncu_pnee.log 2>&1
timeout -k 10 1200 bash scripts/sanitize.sh > $O/r2_sanitizer.log 2>&1
