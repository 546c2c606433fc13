O=gpurun_out
D=$PWD/wasm_pathtracer_b200
timeout -k 5 900 python -m pytest tests -x -q -m gpu -k "scheduling or photon or pnee or config or radiance or extension" > $O/r2k_tests.log 2>&1
echo "exit $?" >> $O/r2k_tests.log
{
python scripts/time_step.py 16 1 > /dev/null
for i in 1 2; do
for v in old new; do
L=$D/libwpt_$v.so; [ $v = new ] && L=$D/libwpt.so
echo -n "$v bvh2 NEE  : "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 16 3 2 1 0
echo -n "$v bvh2 PNEE : "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 16 3 2 2 0
echo -n "$v bvh4 PNEE : "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 16 3 4 2 0
echo -n "$v museum NEE: "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 8 2 2 1 0 0
echo -n "$v museum PNEE: "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 8 2 2 2 0 0
done; done
} > $O/r2k_ab2.log 2>&1
