# N = 8: the bench (weak-scaled headline + strong-scaled target record) and BASELINE config 5 (4K, PNEE + adaptive, 1024 spp budget)
O=gpurun_out
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2b_scale_n8.json 2> $O/r2b_scale_n8.err
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 scripts/run_configs.py 5 > $O/r2b_config5_n8.log 2>&1
