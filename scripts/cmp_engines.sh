python scripts/gpu_first_light.py 0 > gpurun_out/first_light_e0.log 2>&1; echo "parity True/False counts: $(grep -c True gpurun_out/first_light_e0.log) $(grep -c False gpurun_out/first_light_e0.log)"
# warm the GPU, then compare engines back to back (16 spp, mean of 3)
python scripts/time_step.py 16 3 2 1 0 > /dev/null
for e in 0 2 1 0 2; do echo -n "engine=$e: "; python scripts/time_step.py 16 3 2 1 $e; done
for minb in 3 5; do echo -n "engine=0 minb=$minb: "; WPT_MEGA_MINB=$minb python scripts/time_step.py 16 3 2 1 0; done
echo -n "engine=0 bvh4: "; python scripts/time_step.py 16 3 4 1 0
echo -n "engine=0 museum NEE: "; python scripts/time_step.py 4 2 2 1 0 0
