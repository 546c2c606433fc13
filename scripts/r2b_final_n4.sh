O=gpurun_out
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r2b_scale_n4.json 2> $O/r2b_scale_n4.err
timeout -k 10 200 bash scripts/native_dist_check.sh 2 > $O/r2b_native_dist_check_n2.log 2>&1
timeout -k 10 200 bash scripts/native_dist_check.sh 4 > $O/r2b_native_dist_check_n4.log 2>&1
