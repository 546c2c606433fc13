# blocks/SM x vote thresholds on the bench frame (simple BVH2 NEE)
python scripts/time_step.py 16 2 2 1 0 > /dev/null
for mb in 6 7 8; do for th in "20 10" "16 8" "24 12" "12 6"; do set -- $th
echo -n "minb $mb thi $1 tlo $2: "; WPT_MEGA_MINB=$mb WPT_MEGA_THI=$1 WPT_MEGA_TLO=$2 python scripts/time_step.py 16 3 2 1 0; done; done
