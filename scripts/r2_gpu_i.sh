O=gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu --durations=5 > $O/r2i_tests.log 2>&1
echo "exit $?" >> $O/r2i_tests.log
timeout -k 5 600 python bench.py --steps 5 --warmup 3 > $O/r2i_bench.json 2> $O/r2i_bench.err
echo "exit $?" >> $O/r2i_bench.err
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:k_mega -s 1 -c 1 -o $O/r2i_mega_museum -f python scripts/time_step.py 8 1 2 1 0 0 > $O/r2i_ncu_museum.log 2>&1
