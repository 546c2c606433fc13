O=gpurun_out
timeout -k 5 600 python -m pytest tests -x -q -m gpu -k "museum or torus or extension or radiance or features" > $O/r2f_tests.log 2>&1
echo "exit $?" >> $O/r2f_tests.log
AB=$PWD/wasm_pathtracer_b200/libwpt_ab.so
M="8 2 2 1 0 0"
{
python scripts/time_step.py $M > /dev/null
for b in 8 16; do for t in 4 6 8 10 12; do echo -n "v2 MINBG=$b TTORUS=$t: "; WPT_MEGA_MINBG=$b WPT_MEGA_TTORUS=$t timeout -k 5 60 python scripts/time_step.py $M; done; done
for b in 4 5 8; do for t in 6 8; do echo -n "v1 tuning MINBG=$b TTORUS=$t: "; WPT_LIBRARY=$AB WPT_MEGA_MINBG=$b WPT_MEGA_TTORUS=$t timeout -k 5 60 python scripts/time_step.py $M; done; done
for hi in 12 16 24; do for lo in 4 8 12; do echo -n "v2 MINBG=8 TTORUS=8 THI=$hi TLO=$lo: "; WPT_MEGA_MINBG=8 WPT_MEGA_TTORUS=8 WPT_MEGA_THI=$hi WPT_MEGA_TLO=$lo timeout -k 5 60 python scripts/time_step.py $M; done; done
echo -n "v2 PNEE MINBG=8 TTORUS=8: "; WPT_MEGA_MINBG=8 WPT_MEGA_TTORUS=8 timeout -k 5 60 python scripts/time_step.py 8 2 2 2 0 0
echo -n "v2 PNEE MINBG=16 TTORUS=8: "; WPT_MEGA_MINBG=16 WPT_MEGA_TTORUS=8 timeout -k 5 60 python scripts/time_step.py 8 2 2 2 0 0
echo -n "v2 NoNEE MINBG=8 TTORUS=8: "; WPT_MEGA_MINBG=8 WPT_MEGA_TTORUS=8 timeout -k 5 60 python scripts/time_step.py 8 2 2 0 0 0
} > $O/r2f_museum.log 2>&1
