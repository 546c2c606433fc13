# One sweep script for every launch parameter of the path kernels (replaces the per-experiment sweep*.sh / ab*.sh of round 1).
#   bash scripts/sweep.sh VAR "v1 v2 ..." [spp reps bvh type engine scene]
# runs scripts/time_step.py once per value with the environment variable VAR set, e.g.
#   bash scripts/sweep.sh WPT_MEGA_MINB "7 8 9"               # blocks per SM of the triangles / planes BVH2 variant, bench frame
#   bash scripts/sweep.sh WPT_MEGA_MINBG "8 12 16" 8 2 2 1 0 0 # museum (tori): 8 spp, NormalNEE
#   bash scripts/sweep.sh WPT_MEGA_THI "16 20 24"              # traversal-burst threshold
# Variables read by csrc/context.cpp: WPT_MEGA_MINB / MINB4 / MINBG / MINBG4, WPT_MEGA_THI / TLO / TINNER / REPS / CHUNK, WPT_MEGA_TTORUS (lanes parked
# at a torus before the solver phase runs), WPT_MEGA_ZONES ("percent:samples,..." end zones of the slot queue, generic variants), WPT_MEGA_LIST_LEN (samples per
# slot of a strategy round, 0 = segments), WPT_TILE_ORDER (1 = slot order by primary-hit class),
# WPT_WPOOL_CTX / THI / TLO / TSWITCH / REFILL / MINB (engine 4). A/B of two builds: scripts/ab.sh (WPT_LIBRARY).
VAR=$1; VALS=$2; shift 2
ARGS=${@:-16 3 2 1 0}
python scripts/time_step.py $ARGS > /dev/null   # warm-up
for v in $VALS; do
  echo -n "$VAR=$v: "
  env $VAR=$v timeout -k 5 120 python scripts/time_step.py $ARGS
done
