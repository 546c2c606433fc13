# parameter sweep of the warp-pool kernel on the bench frame (bunny 1080p, 16 spp, NormalNEE, BVH2)
for cfg in "128 96 24 16 8" "128 96 28 16 4" "128 64 24 16 8" "128 96 20 16 16" "128 110 30 16 2"; do
  set -- $cfg
  echo -n "ctx $1 thi $2 tlo $3 tswitch $4 refill $5 : "
  WPT_WPOOL_CTX=$1 WPT_WPOOL_THI=$2 WPT_WPOOL_TLO=$3 WPT_WPOOL_TSWITCH=$4 WPT_WPOOL_REFILL=$5 timeout -k 5 60 python scripts/time_step.py 16 3 2 1 0
done
WPT_LIBRARY=$PWD/wasm_pathtracer_b200/libwpt_ab.so WPT_DEBUG_COUNTERS=1 WPT_WPOOL_CTX=128 WPT_WPOOL_THI=96 WPT_WPOOL_TLO=28 WPT_WPOOL_TSWITCH=16 WPT_WPOOL_REFILL=4 timeout -k 5 60 python scripts/time_step.py 16 0 2 1 0 2>&1 | tail -2
