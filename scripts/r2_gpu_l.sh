O=gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu -k "adaptive or random or compute or scheduling or config4 or config3 or tick or dist" > $O/r2l_tests.log 2>&1
echo "exit $?" >> $O/r2l_tests.log
{
for l in 0 4 2 1; do echo "== LIST_LEN=$l full frame"; WPT_MEGA_LIST_LEN=$l timeout -k 5 120 python scripts/target_trace.py | tail -1; done
for l in 0 4 2 1; do echo "== LIST_LEN=$l band of 136 rows (1/8 frame)"; WPT_MEGA_LIST_LEN=$l timeout -k 5 120 python scripts/target_trace.py 472 136 | tail -1; done
for l in 0 4 2 1; do echo "== LIST_LEN=$l band of 272 rows (1/4 frame)"; WPT_MEGA_LIST_LEN=$l timeout -k 5 120 python scripts/target_trace.py 404 272 | tail -1; done
} > $O/r2l_list.log 2>&1
