"""Small fixed workload for ncu: bunny 1080p, NormalNEE, `spp` samples per pixel, `reps` renders."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wasm_pathtracer_b200 as W
from bench import mesh_path, W_, H_
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
bvh = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rtype = int(sys.argv[4]) if len(sys.argv) > 4 else 1
engine = int(sys.argv[5]) if len(sys.argv) > 5 else 0
verts = W.parse_obj(open(mesh_path()).read(), True)
pt = W.PathTracer(W_, H_, W.SCENE_BUNNY, *W.CAM_BUNNY, device=0)
pt.store_mesh(1, verts)
pt.set_config(bvh_kind=bvh, render_type=rtype, engine=engine)
for _ in range(reps):
    pt.reset()
    pt.render_exact(spp)
pt.synchronize()
print(pt.stats())
