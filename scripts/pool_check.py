"""Engine 2 (k_pool) vs engine 0 (k_mega): bit-exact accumulators and counters on small frames, all variants."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import wasm_pathtracer_b200 as W
from bench import mesh_path
ok = True
def bits(a): return np.ascontiguousarray(a).view(np.uint32)
cases = [("bunny bvh2 NEE", 2, 2, 1, 4), ("bunny bvh4 NEE", 2, 4, 1, 4), ("bunny bvh2 PNEE", 2, 2, 2, 4), ("bunny bvh2 NoNEE", 2, 2, 0, 4),
         ("museum bvh2 NEE", 0, 2, 1, 4), ("museum bvh2 PNEE", 0, 2, 2, 4)]
verts = W.parse_obj(open(mesh_path(4)).read(), True)
for name, scene, bvh, rtype, sub in cases:
    for (w, h, spp) in [(128, 72, 3), (333, 211, 2)]:
        cam = W.CAM_BUNNY if scene == 2 else W.CAM_MUSEUM
        a = W.PathTracer(w, h, scene, *cam, device=0)
        if scene == 2: a.store_mesh(1, verts)
        a.set_config(bvh_kind=bvh, render_type=rtype, photon_target=20000, engine=0)
        if rtype == 2: a.build_photons()
        a.reset(); a.render_exact(spp); a.render_exact(1)
        rgb0, c0 = a.accum(); st0 = a.stats()
        a.reset(); a.set_config(engine=2)
        t = time.perf_counter(); a.render_exact(spp); a.render_exact(1); a.synchronize(); dt = time.perf_counter() - t
        rgb1, c1 = a.accum(); st1 = a.stats()
        same = np.array_equal(bits(rgb0), bits(rgb1)) and np.array_equal(c0, c1)
        cnt = tuple(st0[k] == st1[k] for k in ("rays", "node_visits", "paths"))
        print("%-18s %dx%d: accum %s counters %s  (%.1f ms)" % (name, w, h, "OK" if same else "DIFF", cnt, dt * 1e3), flush=True)
        ok = ok and same and all(cnt)
        a.close()
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
