# needs a -DWPT_TUNING build
python scripts/time_step.py 16 1 2 2 0 > /dev/null
for mb in 5 8 12 16; do echo -n "bvh2 PNEE minb $mb: "; WPT_MEGA_MINB=$mb python scripts/time_step.py 16 3 2 2 0; done
for mb in 5 8 12; do echo -n "bvh4 PNEE minb $mb: "; WPT_MEGA_MINB4=$mb python scripts/time_step.py 16 3 4 2 0; done
for mb in 5 8 12; do echo -n "bvh4 NEE minb $mb: "; WPT_MEGA_MINB4=$mb python scripts/time_step.py 16 3 4 1 0; done
for mb in 5 8 12; do echo -n "bvh2 NEE minb $mb: "; WPT_MEGA_MINB=$mb python scripts/time_step.py 16 3 2 1 0; done
for mb in 8 12; do echo -n "bvh2 NoNEE minb $mb: "; WPT_MEGA_MINB=$mb python scripts/time_step.py 16 3 2 0 0; done
