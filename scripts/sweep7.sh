python scripts/time_step.py 16 1 2 1 0 > /dev/null
for mb in 7 8; do for ti in 1 2 3; do echo -n "mega minb $mb tinner $ti: "; WPT_MEGA_MINB=$mb WPT_MEGA_TINNER=$ti python scripts/time_step.py 16 3 2 1 0; done; done
echo -n "mega bvh4: "; python scripts/time_step.py 16 3 4 1 0
echo -n "pool default: "; python scripts/time_step.py 16 3 2 1 2
echo -n "pool tsw32 tlo28: "; WPT_POOL_TSWITCH=32 WPT_POOL_TLO=28 python scripts/time_step.py 16 3 2 1 2
