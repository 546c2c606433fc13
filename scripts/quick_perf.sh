python scripts/time_step.py 16 3 2 1 0 > /dev/null
echo -n "simple  bvh2 NEE : "; python scripts/time_step.py 16 3 2 1 0
echo -n "generic bvh2 NEE : "; WPT_NO_SIMPLE=1 python scripts/time_step.py 16 3 2 1 0
echo -n "simple  bvh4 NEE : "; python scripts/time_step.py 16 3 4 1 0
echo -n "simple  bvh2 PNEE: "; python scripts/time_step.py 16 3 2 2 0
echo -n "simple  bvh2 NoNEE: "; python scripts/time_step.py 16 3 2 0 0
echo -n "museum  bvh2 NEE : "; python scripts/time_step.py 8 2 2 1 0 0
echo -n "museum  bvh2 PNEE: "; python scripts/time_step.py 8 2 2 2 0 0
