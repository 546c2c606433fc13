# round 2, two GPUs: the native NCCL plane against single-session renders (Python host and pure C++ host), then the bench at N = 1 and 2 on the same box
O=gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout -k 10 300 $TR scripts/dist_check.py > $O/r2_dist_check_n$N.log 2>&1
timeout -k 10 300 bash scripts/native_dist_check.sh $N > $O/r2_native_dist_check_n$N.log 2>&1
timeout -k 10 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $O/r2_bench_n1_same_box.json 2> $O/r2_bench_n1_same_box.err
timeout -k 10 300 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err
