# final single-GPU evidence of the round: GPU tests, smoke, bench (both arms), ncu launch list of the bench command, configs 2-4
O=gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu --durations=5 > $O/r2b_tests_full.log 2>&1; echo "exit $?" >> $O/r2b_tests_full.log
timeout -k 5 300 python __graft_entry__.py smoke > $O/r2b_smoke.log 2>&1; echo "exit $?" >> $O/r2b_smoke.log
timeout -k 5 600 python bench.py --steps 10 --warmup 3 > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err; echo "exit $?" >> $O/r2b_bench_n1.err
timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2b_bench_reference_arm.json 2> $O/r2b_bench_reference_arm.err
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-target > $O/r2b_ncu_launch.log 2>&1
timeout -k 5 600 python scripts/run_configs.py 2 3 3m 4 > $O/r2b_configs_n1.log 2>&1
