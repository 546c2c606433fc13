O=gpurun_out
OLD=$PWD/wasm_pathtracer_b200/libwpt_old.so
{
python scripts/time_step.py 16 1 > /dev/null
for g in 0 1 2 4 8; do
echo -n "GUIDE=$g bvh2 NEE 16spp: "; WPT_MEGA_GUIDE=$g timeout -k 5 60 python scripts/time_step.py 16 3 2 1 0
done
for i in 1 2; do
echo -n "old bvh2 NEE : "; WPT_LIBRARY=$OLD timeout -k 5 60 python scripts/time_step.py 16 3 2 1 0
echo -n "new bvh2 NEE : "; timeout -k 5 60 python scripts/time_step.py 16 3 2 1 0
done
for g in 0 2 4; do
echo -n "GUIDE=$g bvh2 NEE 4spp: "; WPT_MEGA_GUIDE=$g timeout -k 5 60 python scripts/time_step.py 4 3 2 1 0
echo -n "GUIDE=$g bvh4 PNEE 8spp: "; WPT_MEGA_GUIDE=$g timeout -k 5 60 python scripts/time_step.py 8 3 4 2 0
echo -n "GUIDE=$g museum: "; WPT_MEGA_GUIDE=$g timeout -k 5 60 python scripts/time_step.py 8 2 2 1 0 0
done
echo -n "old bvh2 PNEE: "; WPT_LIBRARY=$OLD timeout -k 5 60 python scripts/time_step.py 16 3 2 2 0
echo -n "new bvh2 PNEE: "; timeout -k 5 60 python scripts/time_step.py 16 3 2 2 0
echo -n "old bvh4 PNEE: "; WPT_LIBRARY=$OLD timeout -k 5 60 python scripts/time_step.py 16 3 4 2 0
echo -n "new bvh4 PNEE: "; timeout -k 5 60 python scripts/time_step.py 16 3 4 2 0
} > $O/r2g_ab2.log 2>&1
timeout -k 5 200 python scripts/tail_probe.py 2>&1 | grep -v "paths per\|counters" > $O/r2g_tail2.log
timeout -k 5 300 python -m pytest tests -x -q -m gpu -k "parity or features or baseline" > $O/r2g_tests.log 2>&1
