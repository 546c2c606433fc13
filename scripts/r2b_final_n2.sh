O=gpurun_out
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2b_scale_n2.json 2> $O/r2b_scale_n2.err
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 scripts/dist_check.py > $O/r2b_dist_check_n2.log 2>&1
timeout -k 10 200 bash scripts/native_dist_check.sh 2 > $O/r2b_native_dist_check_n2.log 2>&1
