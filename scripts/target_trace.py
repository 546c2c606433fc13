"""The bench's target frame (bunny, BVH4, PNEE, adaptive, 1080p, 64 spp budget) with host wall-clock tracing of the rounds.
usage: python scripts/target_trace.py [region_y region_h]   (a horizontal band of the frame: the per-rank load of a multi-GPU run on one GPU)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wasm_pathtracer_b200 as W
from bench import mesh_path, W_, H_
ry = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rh = int(sys.argv[2]) if len(sys.argv) > 2 else H_
verts = W.parse_obj(open(mesh_path()).read(), True)
tp = W.PathTracer(W_, H_, W.SCENE_BUNNY, *W.CAM_BUNNY, device=0)
tp.store_mesh(1, verts)
tp.set_config(bvh_kind=4, render_type=W.PNEE, photon_target=300000, region_x=0, region_y=ry, region_w=W_, region_h=rh)
t0 = time.perf_counter(); tp.build_photons(); tp.synchronize(); print("photons %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
best = 1e9
for rep in range(4):
    t0 = time.perf_counter(); tp.reset(); tp.synchronize(); t1 = time.perf_counter()
    tp.render_adaptive(W_ * rh * 64); tp.synchronize(); t2 = time.perf_counter()
    best = min(best, t2 - t1)
print("band y=%d h=%d: render_adaptive best of 4: %.2f ms, frame sum %d" % (ry, rh, best * 1e3, int(tp.results(0).astype("uint64").sum())), flush=True)
