for spp in 1 2 4 8 16 32 64; do echo -n "spp=$spp: "; python scripts/time_step.py $spp 2; done
echo -n "engine1 spp=8: "; python scripts/time_step.py 8 2 2 1 1
echo -n "engine1 spp=16: "; python scripts/time_step.py 16 2 2 1 1
