# A/B of two builds of the library on one box: libwpt_old.so (cp of the previous libwpt.so) vs libwpt.so, two rounds per case
D=$PWD/wasm_pathtracer_b200
python scripts/time_step.py 16 1 > /dev/null
for i in 1 2; do
for v in old new; do
L=$D/libwpt_$v.so; [ $v = new ] && L=$D/libwpt.so
echo -n "$v bvh2 NEE   : "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 16 3 2 1 0
echo -n "$v bvh2 PNEE  : "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 16 3 2 2 0
echo -n "$v bvh4 PNEE  : "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 16 3 4 2 0
echo -n "$v museum NEE : "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 8 2 2 1 0 0
echo -n "$v museum PNEE: "; WPT_LIBRARY=$L timeout -k 5 60 python scripts/time_step.py 8 2 2 2 0 0
done; done
