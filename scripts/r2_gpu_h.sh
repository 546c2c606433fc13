O=gpurun_out
{
python scripts/time_step.py 16 1 > /dev/null
for t in 0 1; do for z in "" "10:1" "15:4,10:2,5:1" "20:4,10:2,5:1" "10:4,6:2,4:1" "25:4,12:2,6:1" "20:2,8:1" "30:4,15:2,5:1"; do
echo -n "TILE_ORDER=$t ZONES=$z bvh2 NEE 16spp: "; WPT_TILE_ORDER=$t WPT_MEGA_ZONES=$z timeout -k 5 60 python scripts/time_step.py 16 3 2 1 0
done; done
for z in "" "15:4,10:2,5:1" "25:4,12:2,6:1"; do
echo -n "ZONES=$z bvh2 NEE 4spp: "; WPT_MEGA_ZONES=$z timeout -k 5 60 python scripts/time_step.py 4 3 2 1 0
echo -n "ZONES=$z bvh2 NEE 32spp: "; WPT_MEGA_ZONES=$z timeout -k 5 60 python scripts/time_step.py 32 3 2 1 0
echo -n "ZONES=$z bvh2 PNEE 16spp: "; WPT_MEGA_ZONES=$z timeout -k 5 60 python scripts/time_step.py 16 3 2 2 0
echo -n "ZONES=$z bvh4 PNEE 16spp: "; WPT_MEGA_ZONES=$z timeout -k 5 60 python scripts/time_step.py 16 3 4 2 0
echo -n "ZONES=$z museum TILE_ORDER=0: "; WPT_TILE_ORDER=0 WPT_MEGA_ZONES=$z timeout -k 5 60 python scripts/time_step.py 8 2 2 1 0 0
done
} > $O/r2h_zone2.log 2>&1
WPT_MEGA_ZONES="15:4,10:2,5:1" timeout -k 5 200 python scripts/tail_probe.py 2>&1 | grep -v "counters" > $O/r2h_tail_zone2.log
WPT_MEGA_ZONES="13:4,7:3,5:1" timeout -k 5 300 python -m pytest tests -x -q -m gpu -k "parity or features or baseline or dist or extension" > $O/r2h_tests_zone2.log 2>&1
