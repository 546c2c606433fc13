python scripts/time_step.py 16 1 2 1 0 > /dev/null
for k in 16 8 4 2; do echo -n "segment $k, 16 spp: "; WPT_SEGMENT_LEN_EXPERIMENT=$k python scripts/time_step.py 16 4 2 1 0; done
for k in 16 8 4; do echo -n "segment $k, 16 spp BVH4 PNEE: "; WPT_SEGMENT_LEN_EXPERIMENT=$k python scripts/time_step.py 16 3 4 2 0; done
for k in 16 8; do echo -n "segment $k, museum 8 spp: "; WPT_SEGMENT_LEN_EXPERIMENT=$k python scripts/time_step.py 8 2 2 1 0 0; done
