# checked build (-DWPT_CHECKED) on every device code path incl. the torus phase, end zones and short round slots; then the NCCL warm-up
O=gpurun_out
C=$PWD/wasm_pathtracer_b200/libwpt_checked.so
echo "== checked build (-DWPT_CHECKED: traps on stack overflow, leaf / light / node / slot / pixel / segment-buffer indices out of range)" > $O/r2b_checked.log
WPT_LIBRARY=$C timeout -k 5 200 python scripts/sanitize_case.py 0,1 >> $O/r2b_checked.log 2>&1; echo "== exit code $?" >> $O/r2b_checked.log
WPT_LIBRARY=$C WPT_NO_SIMPLE=1 WPT_MEGA_ZONES="30:4,20:3,10:1" WPT_TILE_ORDER=1 timeout -k 5 200 python scripts/sanitize_case.py 0 >> $O/r2b_checked.log 2>&1; echo "== exit code $? (generic variant, zones + slot order)" >> $O/r2b_checked.log
WPT_LIBRARY=$C timeout -k 5 200 python scripts/time_step.py 16 1 2 1 0 >> $O/r2b_checked.log 2>&1
WPT_LIBRARY=$C timeout -k 5 200 python scripts/time_step.py 8 1 2 1 0 0 >> $O/r2b_checked.log 2>&1
WPT_LIBRARY=$C timeout -k 5 200 python scripts/time_step.py 8 1 2 2 0 0 >> $O/r2b_checked.log 2>&1
WPT_LIBRARY=$C timeout -k 5 200 python scripts/time_step.py 16 1 4 2 0 >> $O/r2b_checked.log 2>&1
WPT_LIBRARY=$C WPT_MEGA_LIST_LEN=2 timeout -k 5 200 python scripts/target_trace.py >> $O/r2b_checked.log 2>&1
WPT_LIBRARY=$C WPT_MEGA_LIST_LEN=1 timeout -k 5 200 python scripts/target_trace.py 472 136 >> $O/r2b_checked.log 2>&1
echo "== unchecked, same box" >> $O/r2b_checked.log
timeout -k 5 200 python scripts/time_step.py 8 2 2 1 0 0 >> $O/r2b_checked.log 2>&1
timeout -k 5 200 python scripts/time_step.py 8 2 2 2 0 0 >> $O/r2b_checked.log 2>&1
timeout -k 5 200 python scripts/time_step.py 16 3 2 1 0 >> $O/r2b_checked.log 2>&1
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2b_warm_n2.json 2> $O/r2b_warm_n2.err
fi
