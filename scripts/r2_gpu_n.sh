O=gpurun_out
D=$PWD/wasm_pathtracer_b200
timeout -k 5 600 python -m pytest tests -x -q -m gpu -k "bvh4 or config3 or primary or incoherent" > $O/r2n_tests.log 2>&1
echo "exit $?" >> $O/r2n_tests.log
{
python scripts/time_step.py 16 1 > /dev/null
for i in 1 2; do
for v in old new; do
L=$D/libwpt_$v.so; [ $v = new ] && L=$D/libwpt.so
echo -n "$v bvh4 PNEE : "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 16 3 4 2 0
echo -n "$v bvh4 NEE  : "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 16 3 4 1 0
echo -n "$v bvh4 NoNEE: "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 16 3 4 0 0
done; done
} > $O/r2n_ab.log 2>&1
