O=gpurun_out
timeout -k 5 900 python -m pytest tests -x -q -m gpu -k "scheduling or photon or pnee or config3 or adaptive or compute or dist" > $O/r2j_tests.log 2>&1
echo "exit $?" >> $O/r2j_tests.log
{
python scripts/time_step.py 16 1 > /dev/null
for i in 1 2; do for e in 0 1; do
echo -n "ENTRY=$e bvh2 PNEE 16spp: "; WPT_PHOTON_ENTRY=$e timeout -k 5 60 python scripts/time_step.py 16 3 2 2 0
echo -n "ENTRY=$e bvh4 PNEE 16spp: "; WPT_PHOTON_ENTRY=$e timeout -k 5 60 python scripts/time_step.py 16 3 4 2 0
done; done
for e in 0 1; do
echo -n "ENTRY=$e museum PNEE 8spp: "; WPT_PHOTON_ENTRY=$e timeout -k 5 60 python scripts/time_step.py 8 2 2 2 0 0
done
} > $O/r2j_entry.log 2>&1
