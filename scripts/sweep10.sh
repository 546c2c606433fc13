python scripts/time_step.py 16 1 2 1 0 > /dev/null
for r in 4 6 8 12 16 32; do echo -n "reps $r: "; WPT_MEGA_REPS=$r python scripts/time_step.py 16 4 2 1 0; done
for r in 4 8; do for ti in 1 3; do echo -n "reps $r tinner $ti: "; WPT_MEGA_REPS=$r WPT_MEGA_TINNER=$ti python scripts/time_step.py 16 4 2 1 0; done; done
for r in 4 8; do for th in "16 8" "24 12" "28 16"; do set -- $th; echo -n "reps $r thi $1 tlo $2: "; WPT_MEGA_REPS=$r WPT_MEGA_THI=$1 WPT_MEGA_TLO=$2 python scripts/time_step.py 16 4 2 1 0; done; done
