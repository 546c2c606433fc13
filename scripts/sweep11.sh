python scripts/time_step.py 16 1 2 1 0 > /dev/null
for mb in 8 9 10 8 9 10; do echo -n "minb $mb: "; WPT_MEGA_MINB=$mb python scripts/time_step.py 16 4 2 1 0; done
for mb in 8 10; do echo -n "minb4 $mb bvh4 pnee: "; WPT_MEGA_MINB4=$mb python scripts/time_step.py 16 3 4 2 0; done
for mb in 8 10; do echo -n "minbg $mb museum: "; WPT_MEGA_MINBG=$mb python scripts/time_step.py 8 2 2 1 0 0; done
