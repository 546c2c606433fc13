# round 2: museum (K_REF variant) blocks/SM sweep + PNEE, and the engine A/B check on the new kernel variants
timeout -k 5 100 python scripts/engine_check.py 0 1 | tail -3
for b in 8 12 16; do
  echo -n "museum NEE  minb $b: "; WPT_MEGA_MINBG=$b timeout -k 5 60 python scripts/time_step.py 8 2 2 1 0 0
  echo -n "museum PNEE minb $b: "; WPT_MEGA_MINBG=$b timeout -k 5 60 python scripts/time_step.py 8 2 2 2 0 0
done
timeout -k 5 200 bash scripts/quick_perf.sh 0
