"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the
oracle on the same seeded inputs. Integer / index work (hit ids, node-visit counts, sample
counts, RGBA8) must be bit-exact; with the shared trig the f32 radiance
accumulators are bit-exact too, so no tolerance is needed anywhere below."""
import os

import numpy as np
import pytest

import oracle_lib as O
import wasm_pathtracer_b200 as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def make_pair(scene, cam, w, h, verts=None, bvh4=False):
    pt = W.PathTracer(w, h, scene, *cam, device=0)
    orc = O.Oracle(w, h, scene, cam)
    if verts is not None:
        pt.store_mesh(1, verts); orc.load_mesh(1, verts)
    if bvh4:
        pt.set_config(bvh_kind=4); orc.rebuild_bvh(True)
    return pt, orc


CASES = [("bunny-meshless", 2, W.CAM_BUNNY, None, False), ("bunny3-bvh2", 2, W.CAM_BUNNY, 3, False), ("bunny4-bvh2", 2, W.CAM_BUNNY, 4, False),
         ("bunny4-bvh4", 2, W.CAM_BUNNY, 4, True), ("museum", 0, W.CAM_MUSEUM, None, False)]


@pytest.mark.parametrize("name,scene,cam,sub,bvh4", CASES)
def test_primary_hits_and_visit_counts_bit_exact(gpu_ok, meshes, name, scene, cam, sub, bvh4):
    pt, orc = make_pair(scene, cam, 256, 144, meshes[sub] if sub else None, bvh4)
    ids, vis, dist = pt.primary_probe()
    oids, ovis, odist = orc.mb_primary_probe()
    assert np.array_equal(ids, oids)
    assert np.array_equal(vis, ovis)
    assert np.array_equal(bits(dist), bits(odist))
    assert (ids >= 0).any()


@pytest.mark.parametrize("name,scene,cam,sub,bvh4", CASES)
def test_incoherent_rays_bit_exact(gpu_ok, meshes, name, scene, cam, sub, bvh4):
    pt, orc = make_pair(scene, cam, 32, 32, meshes[sub] if sub else None, bvh4)
    rng = np.random.default_rng(7)
    n = 30000
    o = (rng.uniform(-4, 4, (n, 3)) + np.array([0, 2, 4])).astype(np.float32)
    d = rng.normal(size=(n, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[:50, 0] = 0.0   # axis-parallel rays: inf / NaN slab arithmetic must match too
    d[50:100, 1] = 0.0
    a = pt.trace_rays(o, d); b = orc.trace_rays(o, d)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    assert np.array_equal(bits(a[1]), bits(b[1]))
    assert np.array_equal(bits(a[3]), bits(b[3]))   # Hit::new-normalised normals


@pytest.mark.parametrize("name,scene,cam,sub,bvh4", CASES)
@pytest.mark.parametrize("rtype", [W.NO_NEE, W.NORMAL_NEE])
def test_radiance_bit_exact(gpu_ok, meshes, name, scene, cam, sub, bvh4, rtype):
    w, h, spp = (160, 90, 3)
    pt, orc = make_pair(scene, cam, w, h, meshes[sub] if sub else None, bvh4)
    pt.set_config(render_type=rtype)
    orc.mb_config(type=rtype, trig=O.TRIG_SHARED)
    pt.render_exact(spp); orc.mb_render_exact(spp, threads=4)
    rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
    assert np.array_equal(cnt, ocnt) and (cnt == spp).all()
    assert np.array_equal(bits(rgb), bits(orgb))
    st, ost = pt.stats(), orc.stats(0)
    assert (st["rays"], st["paths"], st["node_visits"]) == (ost["rays"], ost["paths"], ost["node_visits"])
    assert np.array_equal(pt.results(0), orc.results(0))     # RGBA8: trunc, no gamma, A = 255 (render_target.rs:59-64)


def test_light_debug_mode(gpu_ok, meshes):
    pt, orc = make_pair(2, W.CAM_BUNNY, 96, 64, meshes[3])
    pt.set_config(render_type=W.NORMAL_NEE, light_debug=1)
    orc.mb_config(type=O.NORMAL_NEE, light_debug=True)
    pt.render_exact(2); orc.mb_render_exact(2)
    assert np.array_equal(bits(pt.accum()[0]), bits(orc.accum()[0]))


def test_golden_vectors_through_the_abi(gpu_ok, meshes):
    g = np.load(os.path.join(ROOT, "tests", "golden", "bunny3_48x32.npz"))
    pt = W.PathTracer(48, 32, 2, *W.CAM_BUNNY, device=0)
    pt.store_mesh(1, meshes[3])
    pt.set_config(render_type=W.NORMAL_NEE)
    ids, vis, dist = pt.primary_probe()
    assert np.array_equal(ids, g["ids"]) and np.array_equal(vis, g["visits"]) and np.array_equal(bits(dist), g["dist_bits"])
    pt.render_exact(2)
    assert np.array_equal(bits(pt.accum()[0]), g["rgb_bits"])
    st = pt.stats()
    assert [st["rays"], st["paths"], st["node_visits"]] == g["stats"].tolist()


def test_session_lifecycle_like_the_worker(gpu_ok, meshes):
    """update_camera / update_viewport / update_scene reset accumulation (worker.ts:55-95 call order)."""
    pt, orc = make_pair(2, W.CAM_BUNNY, 64, 48, meshes[3])
    pt.set_config(render_type=W.NORMAL_NEE); orc.mb_config(type=O.NORMAL_NEE)
    pt.render_exact(1)
    pt.update_camera(0.5, 4.0, -1.0, 0.4, 0.2); orc.update_camera(0.5, 4.0, -1.0, 0.4, 0.2)
    assert pt.accum()[1].sum() == 0 and not pt.results(0)[..., :3].any() and (pt.results(0)[..., 3] == 255).all()
    pt.update_viewport(80, 40); orc.update_viewport(80, 40)
    orc.mb_config(type=O.NORMAL_NEE)
    pt.render_exact(2); orc.mb_render_exact(2)
    assert pt.results(0).shape == (40, 80, 4)
    assert np.array_equal(bits(pt.accum()[0]), bits(orc.accum()[0]))
    pt.update_scene(W.SCENE_MUSEUM); orc.update_scene(O.SCENE_MUSEUM)
    pt.render_exact(1); orc.mb_render_exact(1)
    assert np.array_equal(bits(pt.accum()[0]), bits(orc.accum()[0]))


def test_full_size_properties_1080p(gpu_ok, built):
    """BASELINE.json size (1920x1080, stand-in mesh with 81 920 triangles): size-independent properties."""
    verts = W.parse_obj(open(os.path.join(ROOT, "assets", "_gen", "standin_6.obj")).read(), True)
    a = W.PathTracer(1920, 1080, 2, *W.CAM_BUNNY, device=0)
    a.store_mesh(1, verts)
    a.set_config(render_type=W.NORMAL_NEE)
    a.render_exact(4)
    rgb_a, cnt_a = a.accum()
    st_a = a.stats()
    assert (cnt_a == 4).all() and st_a["paths"] == 1920 * 1080 * 4
    # (1) sample streams continue where the last call stopped. Contract B10 sums each call's samples in
    # segments of 8 from +0: 2 + 2 samples are the same samples grouped (c0+c1)+(c2+c3) instead of ((c0+c1)+c2)+c3 —
    # equal up to f32 associativity — while 16 + 16 samples are bit-identical to 32 in one call (same segments).
    a.reset(); a.render_exact(2); a.render_exact(2)
    rgb_b, cnt_b = a.accum()
    assert np.array_equal(cnt_a, cnt_b) and np.allclose(rgb_a, rgb_b, rtol=1e-5, atol=1e-6)
    a.reset(); a.render_exact(32)
    rgb_c, _ = a.accum()
    a.reset(); a.render_exact(16); a.render_exact(16)
    rgb_d, _ = a.accum()
    assert np.array_equal(bits(rgb_c), bits(rgb_d))
    # (2) two band partitions (4-row bands, alternating ranks) == one full-frame render, and the counters add up
    from wasm_pathtracer_b200.dist import rows_of_rank
    parts = []
    tot = {"rays": 0, "node_visits": 0, "paths": 0}
    for r in range(2):
        a.reset(); a.set_config(rank=r, world=2); a.render_exact(4)
        rgb, cnt = a.accum()
        assert (cnt[rows_of_rank(1080, r, 2)] == 4).all() and (cnt[rows_of_rank(1080, 1 - r, 2)] == 0).all()
        parts.append(rgb)
        s = a.stats()
        for k in tot: tot[k] += s[k]
    merged = parts[0].copy(); r1 = rows_of_rank(1080, 1, 2); merged[r1] = parts[1][r1]
    assert np.array_equal(bits(merged), bits(rgb_a))
    assert tot == {k: st_a[k] for k in tot}
    # (3) primary ids: every hit id is a valid shape, visit counts >= 1 (the root guard)
    a.set_config(rank=0, world=1)
    ids, vis, dist = a.primary_probe()
    assert ids.max() < a.scene_info()["num_shapes"] and vis.min() >= 1 and np.isfinite(dist[ids >= 0]).all()
    # (4) a slice of the frame against the oracle: rows 500..519 of the primary probe
    orc = O.Oracle(1920, 1080, 2, O.CAM_BUNNY); orc.load_mesh(1, verts)
    oids, ovis, odist = orc.mb_primary_probe()
    assert np.array_equal(ids, oids) and np.array_equal(vis, ovis) and np.array_equal(bits(dist), bits(odist))


def test_empty_and_degenerate_inputs(gpu_ok, meshes):
    pt = W.PathTracer(1, 1, 2, *W.CAM_BUNNY, device=0)            # smallest viewport
    pt.render_exact(3)
    assert pt.accum()[1].tolist() == [[3]]
    pt.render_exact(0)                                             # zero samples: a no-op
    assert pt.accum()[1].tolist() == [[3]]
    ids, dist, vis, nrm = pt.trace_rays(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert len(ids) == 0
    pt.set_config(region_x=0, region_y=0, region_w=1, region_h=1, rank=1, world=2)   # a rank with no rows
    pt.render_exact(2)
    assert pt.accum()[1].tolist() == [[3]]
    with pytest.raises(W.WptError):
        pt.set_config(region_w=5)                                  # region outside the viewport -> error at render
        pt.render_exact(1)
