# instrumented build (make OUT=../libwpt_ab.so EXTRA=-DWP_INSTR): counters 4.. = runs passes have shade bursts iters run susp stores loads leave_hi leave_dry
for cfg in "96 32 16 16 8" "128 64 8 16 8" "128 96 8 16 8"; do
  set -- $cfg
  echo "ctx $1 thi $2 tlo $3 tswitch $4 refill $5"
  WPT_LIBRARY=$PWD/wasm_pathtracer_b200/libwpt_ab.so WPT_DEBUG_COUNTERS=1 WPT_WPOOL_CTX=$1 WPT_WPOOL_THI=$2 WPT_WPOOL_TLO=$3 WPT_WPOOL_TSWITCH=$4 WPT_WPOOL_REFILL=$5 timeout -k 5 60 python scripts/time_step.py 16 0 2 1 0 2>&1 | tail -2
done
