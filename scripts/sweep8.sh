python scripts/time_step.py 16 1 4 1 0 > /dev/null
for mb in 4 5 6 7 8; do echo -n "bvh4 simple NEE minb $mb: "; WPT_MEGA_MINB4=$mb python scripts/time_step.py 16 3 4 1 0; done
for mb in 4 6 8; do echo -n "bvh4 simple PNEE minb $mb: "; WPT_MEGA_MINB4=$mb python scripts/time_step.py 16 3 4 2 0; done
for mb in 4 5 6 7 8; do echo -n "museum NEE minb $mb: "; WPT_MEGA_MINBG=$mb python scripts/time_step.py 8 2 2 1 0 0; done
for mb in 4 6 8; do echo -n "museum PNEE minb $mb: "; WPT_MEGA_MINBG=$mb python scripts/time_step.py 8 2 2 2 0 0; done
