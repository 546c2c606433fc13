"""The bench's target frame (bunny, BVH4, PNEE, adaptive, 1080p, 64 spp budget) with host wall-clock tracing of the rounds."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wasm_pathtracer_b200 as W
from bench import mesh_path, W_, H_
verts = W.parse_obj(open(mesh_path()).read(), True)
tp = W.PathTracer(W_, H_, W.SCENE_BUNNY, *W.CAM_BUNNY, device=0)
tp.store_mesh(1, verts)
tp.set_config(bvh_kind=4, render_type=W.PNEE, photon_target=300000)
t0 = time.perf_counter(); tp.build_photons(); tp.synchronize(); print("photons %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
for rep in range(3):
    t0 = time.perf_counter(); tp.reset(); tp.synchronize(); t1 = time.perf_counter()
    tp.render_adaptive(W_ * H_ * 64); tp.synchronize(); t2 = time.perf_counter()
    print("rep %d: reset %.2f ms, render_adaptive %.2f ms" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
