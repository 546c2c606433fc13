# round 2: the bench at N = 1, 2, 4, 8 on one box (weak-scaled headline + strong-scaled target record in every line)
O=gpurun_out
MAXN=${1:-8}
timeout -k 10 300 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r2_scale_n1.json 2> $O/r2_scale_n1.err
for N in 2 4 8; do
  [ $N -le $MAXN ] || continue
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 10 --warmup 3 > $O/r2_scale_n$N.json 2> $O/r2_scale_n$N.err
done
timeout -k 10 200 bash scripts/native_dist_check.sh 2 > $O/r2_native_dist_check_n2.log 2>&1
