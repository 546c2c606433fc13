O=gpurun_out
echo "== checked build (-DWPT_CHECKED: traps on stack overflow, leaf / light / node / slot / pixel indices out of range)" > $O/r2_checked.log
WPT_LIBRARY=$PWD/wasm_pathtracer_b200/libwpt_checked.so timeout -k 5 200 python scripts/sanitize_case.py 0,1 >> $O/r2_checked.log 2>&1
echo "== exit code $?" >> $O/r2_checked.log
WPT_LIBRARY=$PWD/wasm_pathtracer_b200/libwpt_checked.so timeout -k 5 200 python scripts/time_step.py 16 1 2 1 0 >> $O/r2_checked.log 2>&1
WPT_LIBRARY=$PWD/wasm_pathtracer_b200/libwpt_checked.so timeout -k 5 200 python scripts/time_step.py 8 1 2 2 0 0 >> $O/r2_checked.log 2>&1
WPT_LIBRARY=$PWD/wasm_pathtracer_b200/libwpt_checked.so timeout -k 5 200 python scripts/time_step.py 16 1 4 2 0 >> $O/r2_checked.log 2>&1
timeout -k 5 200 python scripts/tail_probe.py > $O/r2_tail_hist.log 2>&1
