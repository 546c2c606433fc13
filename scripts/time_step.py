import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wasm_pathtracer_b200 as W
from bench import mesh_path, W_, H_
spp = int(sys.argv[1]); reps = int(sys.argv[2]); bvh = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rtype = int(sys.argv[4]) if len(sys.argv) > 4 else 1; engine = int(sys.argv[5]) if len(sys.argv) > 5 else 0
scene = int(sys.argv[6]) if len(sys.argv) > 6 else 2
verts = W.parse_obj(open(mesh_path()).read(), True)
cam = W.CAM_BUNNY if scene == 2 else W.CAM_MUSEUM
pt = W.PathTracer(W_, H_, scene, *cam, device=0)
if scene == 2: pt.store_mesh(1, verts)
pt.set_config(bvh_kind=bvh, render_type=rtype, engine=engine)
best = 1e9
for _ in range(reps + 1):
    pt.reset(); pt.synchronize()
    t = time.perf_counter(); pt.render_exact(spp); pt.synchronize(); dt = time.perf_counter() - t
    best = min(best, dt)
st = pt.stats()
print("%.2f ms  %.0f Mrays/s  (rays %d)" % (best * 1e3, st["rays"] / best / 1e6, st["rays"]))
