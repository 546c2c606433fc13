# launch time of the path kernel against samples per pixel (1 segment per pixel up to 8 spp): the intercept is the fixed cost of a launch (ramp-up + tail)
for cfg in "4 2" "2 1"; do set -- $cfg
for spp in 1 2 4 8 16 32; do echo -n "bvh$1 type $2 spp $spp: "; timeout -k 5 60 python scripts/time_step.py $spp 3 $1 $2 0; done; done
