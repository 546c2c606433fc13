python scripts/gpu_first_light.py 2 > gpurun_out/first_light_e2.log 2>&1; echo "parity True/False counts: $(grep -c True gpurun_out/first_light_e2.log) $(grep -c False gpurun_out/first_light_e2.log)"
python scripts/time_step.py 16 3 2 1 2 > /dev/null
for ti in 1 2 4 8 100; do for th in "16 8" "24 12" "12 4"; do set -- $th
echo -n "engine=2 tinner=$ti thi=$1 tlo=$2: "; WPT_MEGA_TINNER=$ti WPT_MEGA_THI=$1 WPT_MEGA_TLO=$2 python scripts/time_step.py 16 3 2 1 2; done; done
echo -n "engine=0: "; python scripts/time_step.py 16 3 2 1 0
