# compute-sanitizer on the small cases of scripts/sanitize_case.py: memcheck on all engines, racecheck + synccheck on
# the persistent kernel (warp-synchronous code: ballots, shuffles, redux).
# usage: bash scripts/sanitize.sh > profiles/r2_sanitizer.log     (on a GPU box)
for tool in memcheck racecheck synccheck; do
  echo "==== compute-sanitizer --tool $tool"
  timeout -k 10 900 compute-sanitizer --tool $tool --error-exitcode 1 python scripts/sanitize_case.py 0,1 2>&1 | grep -v "^$" | tail -25
  echo "==== exit code $?"
done
