# A/B of two builds of the library on the bench frame: libwpt.so vs libwpt_ab.so (make OUT=../libwpt_ab.so EXTRA=-D...)
AB=$PWD/wasm_pathtracer_b200/libwpt_ab.so
timeout -k 5 100 python scripts/engine_check.py 0 1 | tail -2
WPT_LIBRARY=$AB timeout -k 5 100 python scripts/engine_check.py 0 1 | tail -2
for i in 1 2; do
  echo -n "A bvh2 NEE : "; timeout -k 5 60 python scripts/time_step.py 16 3 2 1 0
  echo -n "B bvh2 NEE : "; WPT_LIBRARY=$AB timeout -k 5 60 python scripts/time_step.py 16 3 2 1 0
done
echo -n "A bvh2 PNEE: "; timeout -k 5 60 python scripts/time_step.py 16 3 2 2 0
echo -n "B bvh2 PNEE: "; WPT_LIBRARY=$AB timeout -k 5 60 python scripts/time_step.py 16 3 2 2 0
echo -n "A bvh4 NEE : "; timeout -k 5 60 python scripts/time_step.py 16 3 4 1 0
echo -n "B bvh4 NEE : "; WPT_LIBRARY=$AB timeout -k 5 60 python scripts/time_step.py 16 3 4 1 0
echo -n "A museum   : "; timeout -k 5 60 python scripts/time_step.py 8 2 2 1 0 0
echo -n "B museum   : "; WPT_LIBRARY=$AB timeout -k 5 60 python scripts/time_step.py 8 2 2 1 0 0
