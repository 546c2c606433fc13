timeout -k 10 1500 bash scripts/sanitize.sh > gpurun_out/r2_sanitizer.log 2>&1
