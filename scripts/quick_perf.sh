# one line per kernel variant: best of 3 renders of the 1080p frame (engine = $1, default 0)
E=${1:-0}
python scripts/time_step.py 16 3 2 1 $E > /dev/null
echo -n "simple  bvh2 NEE : "; python scripts/time_step.py 16 3 2 1 $E
echo -n "simple  bvh4 NEE : "; python scripts/time_step.py 16 3 4 1 $E
echo -n "simple  bvh2 PNEE: "; python scripts/time_step.py 16 3 2 2 $E
echo -n "simple  bvh4 PNEE: "; python scripts/time_step.py 16 3 4 2 $E
echo -n "simple  bvh2 NoNEE: "; python scripts/time_step.py 16 3 2 0 $E
echo -n "museum  bvh2 NEE : "; python scripts/time_step.py 8 2 2 1 $E 0
echo -n "museum  bvh2 PNEE: "; python scripts/time_step.py 8 2 2 2 $E 0
