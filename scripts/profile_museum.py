"""Small fixed museum workload for ncu: 1080p, NormalNEE, `spp` samples."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wasm_pathtracer_b200 as W
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rtype = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pt = W.PathTracer(1920, 1080, W.SCENE_MUSEUM, *W.CAM_MUSEUM, device=0)
pt.set_config(render_type=rtype)
for _ in range(reps):
    pt.reset(); pt.render_exact(spp)
pt.synchronize()
print(pt.stats())
