python scripts/time_step.py 16 1 2 2 0 > /dev/null
for i in 1 2; do
echo -n "A bvh2 PNEE: "; python scripts/time_step.py 16 3 2 2 0
echo -n "B bvh2 PNEE: "; WPT_LIBRARY=$PWD/wasm_pathtracer_b200/libwpt_ab.so python scripts/time_step.py 16 3 2 2 0
done
echo -n "A bvh2 NEE: "; python scripts/time_step.py 16 3 2 1 0
echo -n "B bvh2 NEE: "; WPT_LIBRARY=$PWD/wasm_pathtracer_b200/libwpt_ab.so python scripts/time_step.py 16 3 2 1 0
echo -n "A bvh4 PNEE: "; python scripts/time_step.py 16 3 4 2 0
echo -n "B bvh4 PNEE: "; WPT_LIBRARY=$PWD/wasm_pathtracer_b200/libwpt_ab.so python scripts/time_step.py 16 3 4 2 0
