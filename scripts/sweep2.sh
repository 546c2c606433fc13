python scripts/time_step.py 16 3 2 1 0 > /dev/null
for minb in 4 5 6; do for th in "16 8" "12 6" "20 10" "24 16" "8 4"; do for ti in 2 4; do set -- $th
echo -n "minb=$minb thi=$1 tlo=$2 ti=$ti: "; WPT_MEGA_MINB=$minb WPT_MEGA_THI=$1 WPT_MEGA_TLO=$2 WPT_MEGA_TINNER=$ti python scripts/time_step.py 16 3 2 1 0; done; done; done
