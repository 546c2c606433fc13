python scripts/time_step.py 16 3 2 1 0 > /dev/null
for minb in 6 7 8 6; do echo -n "minb=$minb: "; WPT_MEGA_MINB=$minb WPT_MEGA_THI=20 WPT_MEGA_TLO=10 WPT_MEGA_TINNER=2 python scripts/time_step.py 16 4 2 1 0; done
