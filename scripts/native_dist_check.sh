# Native multi-GPU plane without Python: N processes of tools/wpt_render, one per GPU, must print the frame hash of the 1-GPU run.
# usage: bash scripts/native_dist_check.sh [N=2]     (run on a box with >= N GPUs, e.g. gpurun --gpus 2)
N=${1:-2}
OBJ=assets/_gen/standin_4.obj
for mode in "--type 1" "--type 2 --photons 20000" "--type 1 --adaptive" "--type 2 --bvh 4 --adaptive --photons 20000"; do
  rm -f /tmp/wpt_nccl_id
  ref=$(tools/wpt_render --quiet --obj $OBJ --size 320x181 --frame-spp 11 $mode --world 1 --rank 0 --device 0 --out "" | grep frame_fnv1a | sed 's/.*frame_fnv1a": "\([0-9a-f]*\)".*/\1/')
  pids=""
  for r in $(seq 0 $((N - 1))); do
    tools/wpt_render --quiet --obj $OBJ --size 320x181 --frame-spp 11 $mode --world $N --rank $r --device $r --nccl-id-file /tmp/wpt_nccl_id --out "" > /tmp/wpt_rank_$r.json &
    pids="$pids $!"
  done
  for p in $pids; do wait $p; done
  for r in $(seq 0 $((N - 1))); do
    got=$(grep frame_fnv1a /tmp/wpt_rank_$r.json | sed 's/.*frame_fnv1a": "\([0-9a-f]*\)".*/\1/')
    echo "mode [$mode] rank $r of $N: frame hash $got, 1-GPU hash $ref: $([ "$got" = "$ref" ] && echo SAME || echo DIFFERENT)"
  done
done
