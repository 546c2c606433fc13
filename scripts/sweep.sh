#!/bin/bash
# sweep megakernel launch parameters: prints ms for bunny 1080p at 8 spp
for minb in 3 4 5 6; do for th in "20 10" "16 8" "24 12" "28 16" "12 4" "8 2" "32 16"; do set -- $th
  echo -n "minb=$minb thi=$1 tlo=$2: "
  WPT_MEGA_MINB=$minb WPT_MEGA_THI=$1 WPT_MEGA_TLO=$2 python scripts/time_step.py 8 3
done; done
