O=gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu -k "adaptive or compute or config4 or config3 or tick or error" > $O/r2m_tests.log 2>&1
echo "exit $?" >> $O/r2m_tests.log
timeout -k 5 120 python scripts/target_trace.py > $O/r2m_target.log 2>&1
timeout -k 5 400 python bench.py --steps 5 --warmup 3 > $O/r2m_bench.json 2> $O/r2m_bench.err
