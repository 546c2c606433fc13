"""First-light GPU check: primary probes, ray batches and small renders vs the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import wasm_pathtracer_b200 as W
import oracle_lib as O

def bits(a): return np.ascontiguousarray(a).view(np.uint32)

ENGINE = int(sys.argv[1]) if len(sys.argv) > 1 else 0

def check_scene(name, scene, cam, verts, w, h, bvh4, types, spp):
    print("==", name, "bvh4" if bvh4 else "bvh2", w, h)
    pt = W.PathTracer(w, h, scene, *cam, device=0)
    pt.set_config(engine=ENGINE)
    orc = O.Oracle(w, h, scene, cam)
    if verts is not None:
        pt.store_mesh(1, verts); orc.load_mesh(1, verts)
    if bvh4:
        pt.set_config(bvh_kind=4); orc.rebuild_bvh(True)
    t = time.time(); ids, vis, dist = pt.primary_probe(); tg = time.time() - t
    t = time.time(); oids, ovis, odist = orc.mb_primary_probe(); to = time.time() - t
    print("  probe ids eq", np.array_equal(ids, oids), "visits eq", np.array_equal(vis, ovis), "dist eq", np.array_equal(bits(dist), bits(odist)), "gpu %.3fs cpu %.3fs" % (tg, to))
    if not np.array_equal(ids, oids): print("   id mismatches", (ids != oids).sum(), np.argwhere(ids != oids)[:5])
    if not np.array_equal(vis, ovis): print("   visit mismatches", (vis != ovis).sum(), np.argwhere(vis != ovis)[:5], vis[vis != ovis][:5], ovis[vis != ovis][:5])
    rng = np.random.default_rng(1)
    n = 20000
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32) + np.array([0, 2, 5], np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    a = pt.trace_rays(o, d); b = orc.trace_rays(o, d)
    print("  rays ids", np.array_equal(a[0], b[0]), "dist", np.array_equal(bits(a[1]), bits(b[1])), "visits", np.array_equal(a[2], b[2]), "normals", np.array_equal(bits(a[3]), bits(b[3])), "hit frac", (a[0] >= 0).mean())
    if not np.array_equal(bits(a[3]), bits(b[3])):
        bad = np.argwhere((bits(a[3]) != bits(b[3])).any(1))[:5, 0]; print("   normal mismatch", len(bad), a[3][bad], b[3][bad], a[0][bad])
    for ty in types:
        pt.set_config(render_type=ty); pt.reset()
        orc.mb_config(type=ty, trig=O.TRIG_SHARED); orc.reset()
        t = time.time(); pt.render_exact(spp); pt.synchronize(); tg = time.time() - t
        t = time.time(); orc.mb_render_exact(spp, threads=8); to = time.time() - t
        rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
        st = pt.stats(); ost = orc.stats(0)
        eq = np.array_equal(bits(rgb), bits(orgb))
        print("  render type", ty, "acc eq", eq, "cnt eq", np.array_equal(cnt, ocnt), "rays", st["rays"], ost["rays"], "visits", st["node_visits"], ost["node_visits"], "paths", st["paths"], ost["paths"], "iters", st["iterations"], "gpu %.3fs cpu %.3fs" % (tg, to))
        if not eq:
            bad = (bits(rgb) != bits(orgb)).any(2); print("   bad pixels", bad.sum(), "of", bad.size, "max abs", np.nanmax(np.abs(rgb - orgb)))
        print("  rgba eq", np.array_equal(pt.results(0), orc.results(0)))
    pt.close(); orc.close()

gen = os.path.join(ROOT, "assets", "_gen")
v3 = W.parse_obj(open(os.path.join(gen, "standin_3.obj")).read(), True)
v4 = W.parse_obj(open(os.path.join(gen, "standin_4.obj")).read(), True)
check_scene("bunny-meshless", 2, W.CAM_BUNNY, None, 128, 96, False, [0, 1], 2)
check_scene("bunny-3", 2, W.CAM_BUNNY, v3, 160, 90, False, [0, 1], 4)
check_scene("bunny-4", 2, W.CAM_BUNNY, v4, 320, 180, False, [1], 4)
check_scene("bunny-4", 2, W.CAM_BUNNY, v4, 320, 180, True, [1], 2)
check_scene("museum", 0, W.CAM_MUSEUM, None, 192, 128, False, [0, 1], 2)
print("done")
