"""Timeline of one k_mega launch (needs the -DMEGA_INSTR build as libwpt_ab.so): first start, first "queue empty", last exit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["WPT_LIBRARY"] = os.path.join(ROOT, "wasm_pathtracer_b200", "libwpt_ab.so")
import numpy as np, ctypes as C
import wasm_pathtracer_b200 as W
from bench import mesh_path, W_, H_
verts = W.parse_obj(open(mesh_path()).read(), True)
for bvh, rtype, spp in ((2, 1, 16), (2, 1, 4), (4, 2, 8), (4, 2, 1)):
    pt = W.PathTracer(W_, H_, W.SCENE_BUNNY, *W.CAM_BUNNY, device=0)
    pt.store_mesh(1, verts)
    pt.set_config(bvh_kind=bvh, render_type=rtype)
    if rtype == 2: pt.build_photons()
    for rep in range(2):
        pt.reset(); pt.render_exact(spp); pt.synchronize()
    os.environ["WPT_DEBUG_COUNTERS"] = "1"
    sys.stderr.write("bvh%d type %d spp %d: " % (bvh, rtype, spp)); sys.stderr.flush()
    pt.stats()
    del os.environ["WPT_DEBUG_COUNTERS"]
    pt.close()
