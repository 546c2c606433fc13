python scripts/time_step.py 16 3 2 1 0 > /dev/null
for th in "20 10 2" "16 8 2" "24 12 2" "12 6 2" "20 6 2" "20 14 2" "16 4 2" "20 10 1" "20 10 4"; do set -- $th; echo -n "thi=$1 tlo=$2 ti=$3: "; WPT_MEGA_THI=$1 WPT_MEGA_TLO=$2 WPT_MEGA_TINNER=$3 python scripts/time_step.py 16 4 2 1 0; done
for minb in 6 8; do echo -n "minb=$minb: "; WPT_MEGA_MINB=$minb python scripts/time_step.py 16 4 2 1 0; done
echo -n "bvh4: "; python scripts/time_step.py 16 4 4 1 0
echo -n "PNEE: "; python scripts/time_step.py 16 4 2 2 0
echo -n "museum NEE: "; python scripts/time_step.py 8 2 2 1 0 0
echo -n "spp64: "; python scripts/time_step.py 64 2 2 1 0
