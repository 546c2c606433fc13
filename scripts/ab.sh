# A/B of two builds in the same call: libwpt.so (A) vs libwpt_ab.so (B), alternating
python scripts/time_step.py 16 1 2 1 0 > /dev/null
for i in 1 2 3; do
echo -n "A: "; python scripts/time_step.py 16 4 2 1 0
echo -n "B: "; WPT_LIBRARY=$PWD/wasm_pathtracer_b200/libwpt_ab.so python scripts/time_step.py 16 4 2 1 0
done
