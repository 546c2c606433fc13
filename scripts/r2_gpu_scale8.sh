# N = 8 only (and N = 1 on the same box for the ratio)
O=gpurun_out
timeout -k 10 300 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r2b_scale_n1.json 2> $O/r2b_scale_n1.err
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2b_scale_n8.json 2> $O/r2b_scale_n8.err
