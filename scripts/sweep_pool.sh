for cfg in "4 384 20 24 16" "4 384 20 24 32" "4 384 28 24 32" "4 384 24 20 24"; do set -- $cfg
echo "minb $1 slots $2 tlo $3 thi $4 tsw $5: "; WPT_DEBUG_COUNTERS=1 WPT_POOL_MINB=$1 WPT_POOL_SLOTS=$2 WPT_POOL_TLO=$3 WPT_POOL_THI=$4 WPT_POOL_TSWITCH=$5 timeout 60 python scripts/time_step.py 16 2 2 1 2 2>&1 | tail -2; done
