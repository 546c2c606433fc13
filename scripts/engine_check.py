"""Quick A/B of two engines on one GPU: bit-equality of a small and a ragged render, then timings of the 1080p frame.
usage: python scripts/engine_check.py [engine_a] [engine_b]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import wasm_pathtracer_b200 as W
from bench import mesh_path
ea = int(sys.argv[1]) if len(sys.argv) > 1 else 0
eb = int(sys.argv[2]) if len(sys.argv) > 2 else 3
v4 = W.parse_obj(open(mesh_path(4)).read(), True)
ok = True
for (w, h, scene, bvh, rtype, spp) in [(128, 72, 2, 2, 1, 3), (333, 211, 2, 2, 2, 9), (97, 61, 2, 4, 1, 17), (160, 90, 0, 2, 1, 5), (160, 90, 0, 2, 2, 5), (640, 360, 2, 4, 2, 16)]:
    pt = W.PathTracer(w, h, scene, *(W.CAM_BUNNY if scene == 2 else W.CAM_MUSEUM), device=0)
    if scene == 2: pt.store_mesh(1, v4)
    pt.set_config(bvh_kind=bvh, render_type=rtype, photon_target=30000, engine=ea)
    if rtype == 2: pt.build_photons()
    res = []
    for e in (ea, eb):
        pt.reset(); pt.set_config(engine=e); pt.render_exact(spp); pt.render_exact(2)
        rgb, cnt = pt.accum(); st = pt.stats()
        res.append((rgb.view(np.uint32).copy(), cnt.copy(), (st["rays"], st["paths"], st["node_visits"])))
    same = np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]
    ok &= same
    print("%4dx%-4d scene %d bvh%d type %d spp %2d: %s %s %s" % (w, h, scene, bvh, rtype, spp, "same" if same else "DIFFERENT", res[0][2], res[1][2]), flush=True)
    pt.close()
print("ALL SAME" if ok else "MISMATCH", flush=True)
