"""Small renders of every device code path for compute-sanitizer (scripts/sanitize.sh): both scenes, BVH2 / BVH4,
NoNEE / NEE / PNEE (photon warm-up + octree build), adaptive and random strategy rounds, both engines."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import wasm_pathtracer_b200 as W
from bench import mesh_path
v = W.parse_obj(open(mesh_path(3)).read(), True)
engines = [int(e) for e in (sys.argv[1] if len(sys.argv) > 1 else "0,1").split(",")]
for engine in engines:
    for scene, cam, bvh, rtype in ((2, W.CAM_BUNNY, 2, W.NORMAL_NEE), (2, W.CAM_BUNNY, 4, W.PNEE), (0, W.CAM_MUSEUM, 2, W.PNEE), (0, W.CAM_MUSEUM, 2, W.NO_NEE)):
        pt = W.PathTracer(64, 44, scene, *cam, device=0)
        if scene == 2:
            pt.store_mesh(1, v)
        pt.set_config(bvh_kind=bvh, render_type=rtype, photon_target=3000, engine=engine)
        if rtype == W.PNEE:
            pt.build_photons()
        pt.render_exact(3)
        pt.render_adaptive(64 * 44 * 9 + 17)
        pt.render_random(5000)
        pt.primary_probe()
        img = pt.results(0)
        st = pt.stats()
        print("engine %d scene %d bvh%d type %d: %d rays, frame sum %d" % (engine, scene, bvh, rtype, st["rays"], int(img.astype(np.uint64).sum())), flush=True)
        pt.close()
print("done")
