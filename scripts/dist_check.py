"""Run under torchrun: band-partitioned render + the native NCCL all-gather must equal a single-session render bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import wasm_pathtracer_b200 as W
from wasm_pathtracer_b200.dist import attach

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
verts = W.parse_obj(open(os.path.join(ROOT, "assets", "_gen", "standin_4.obj")).read(), True)
w, h = 320, 181   # odd height: ragged partition
ok = True
for rtype, mode in ((W.NORMAL_NEE, "exact"), (W.PNEE, "exact"), (W.NORMAL_NEE, "adaptive"), (W.PNEE, "adaptive")):
    pt = W.PathTracer(w, h, W.SCENE_BUNNY, *W.CAM_BUNNY, device=local)
    pt.store_mesh(1, verts)
    pt.set_config(render_type=rtype, photon_target=20000)
    attach(pt, rank, world)     # band partition + the library's native NCCL plane: accumulator all-gather, photon shots split over ranks (uint32 sum-allreduce)
    if mode == "exact":
        pt.render_exact(3)
        pt.gather_frame()
    else:
        pt.render_adaptive(w * h * 11 + 123)
    rgb, cnt = pt.accum()
    ref = W.PathTracer(w, h, W.SCENE_BUNNY, *W.CAM_BUNNY, device=local)
    ref.store_mesh(1, verts)
    ref.set_config(render_type=rtype, photon_target=20000)
    if mode == "exact": ref.render_exact(3)
    else: ref.render_adaptive(w * h * 11 + 123)
    rrgb, rcnt = ref.accum()
    if rtype == W.PNEE:
        a, b = pt.photons(), ref.photons()
        tree_same = a[3] == b[3] and all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip(a[:3], b[:3]))
        print("rank %d photons split over %d ranks == single session: %s (%d photons, %d shots)" % (rank, world, tree_same, len(a[0]), a[3]), flush=True)
        ok = ok and tree_same
    same = np.array_equal(cnt, rcnt) and np.array_equal(rgb.view(np.uint32), rrgb.view(np.uint32)) and np.array_equal(pt.results(0), ref.results(0))
    print("rank %d type %d %s: partitioned == single session: %s" % (rank, rtype, mode, same), flush=True)
    ok = ok and same
    pt.detach_nccl()
    pt.close(); ref.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
