"""GPU vs oracle at the BASELINE.json configurations themselves (VERDICT r1 "missing" #3): the same bit-exact gates as
tests/test_gpu_parity.py / test_gpu_features.py, at full size, with the oracle on all host threads.

  config 2: bunny (stand-in mesh, 81 920 triangles), BVH2, 1920x1080, 16 spp, NormalNEE — radiance bits of the whole frame
  config 3: bunny, BVH4 collapse AND photon-based NEE together, 300 000 photons, 480x270 (and the museum: 108 lights)
  config 4: museum (tori), 3840x2160, adaptive strategy over the full frame, budget reaching into the second adaptive round
"""
import os

import numpy as np
import pytest

import oracle_lib as O
import wasm_pathtracer_b200 as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THREADS = os.cpu_count() or 1


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def mesh6(built):
    return W.parse_obj(open(os.path.join(ROOT, "assets", "_gen", "standin_6.obj")).read(), True)


def test_config2_full_frame_radiance_bits(gpu_ok, mesh6):
    w, h, spp = 1920, 1080, 16
    pt = W.PathTracer(w, h, 2, *W.CAM_BUNNY, device=0); pt.store_mesh(1, mesh6)
    orc = O.Oracle(w, h, 2, O.CAM_BUNNY); orc.load_mesh(1, mesh6)
    pt.set_config(render_type=W.NORMAL_NEE); orc.mb_config(type=O.NORMAL_NEE, trig=O.TRIG_SHARED)
    pt.render_exact(spp); orc.mb_render_exact(spp, threads=THREADS)
    rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
    assert (cnt == spp).all() and np.array_equal(cnt, ocnt)
    assert np.array_equal(bits(rgb), bits(orgb))
    st, ost = pt.stats(), orc.stats(0)
    assert (st["rays"], st["paths"], st["node_visits"]) == (ost["rays"], ost["paths"], ost["node_visits"])
    assert np.array_equal(pt.results(0), orc.results(0))
    ids, vis, dist = pt.primary_probe(); oids, ovis, odist = orc.mb_primary_probe()
    assert np.array_equal(ids, oids) and np.array_equal(vis, ovis) and np.array_equal(bits(dist), bits(odist))


@pytest.mark.parametrize("scene,cam,bvh4", [(2, W.CAM_BUNNY, True), (0, W.CAM_MUSEUM, False)])
def test_config3_bvh4_with_pnee_300k_photons(gpu_ok, mesh6, scene, cam, bvh4):
    """BVH4 traversal and the photon-tree light choice in one render (the museum cannot be collapsed: it has a
    17-shape leaf, bvh4.rs:22,135 — there PNEE runs over its 108 lights on BVH2)."""
    w, h, spp = 480, 270, 4
    pt = W.PathTracer(w, h, scene, *cam, device=0)
    orc = O.Oracle(w, h, scene, cam)
    if scene == 2:
        pt.store_mesh(1, mesh6); orc.load_mesh(1, mesh6)
    if bvh4:
        pt.set_config(bvh_kind=4); orc.rebuild_bvh(True)
    pt.set_config(render_type=W.PNEE, photon_target=300000); orc.mb_config(type=O.PNEE, photon_target=300000, trig=O.TRIG_SHARED)
    pt.build_photons(); orc.mb_build_photons(threads=THREADS)
    light, loc, wgt, shots = pt.photons(); olight, oloc, owgt, oshots = orc.mb_photons()
    assert len(light) == 300000 and shots == oshots
    assert np.array_equal(light, olight) and np.array_equal(bits(loc), bits(oloc)) and np.array_equal(bits(wgt), bits(owgt))
    meta, cum, bins_ = pt.photon_tree(); ometa, ocum, obins = orc.mb_photon_tree()
    assert np.array_equal(meta, ometa) and np.array_equal(bits(cum), bits(ocum)) and np.array_equal(bits(bins_), bits(obins))
    pt.render_exact(spp); orc.mb_render_exact(spp, threads=THREADS)
    rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
    assert np.array_equal(cnt, ocnt) and np.array_equal(bits(rgb), bits(orgb))
    st, ost = pt.stats(), orc.stats(0)
    assert (st["rays"], st["paths"], st["node_visits"]) == (ost["rays"], ost["paths"], ost["node_visits"])
    if bvh4:
        ids, vis, dist = pt.primary_probe(); oids, ovis, odist = orc.mb_primary_probe()
        assert np.array_equal(ids, oids) and np.array_equal(vis, ovis) and np.array_equal(bits(dist), bits(odist))


def test_config4_museum_adaptive_4k_two_rounds(gpu_ok):
    w, h = 3840, 2160
    n = w * h
    pt = W.PathTracer(w, h, 0, *W.CAM_MUSEUM, device=0)
    orc = O.Oracle(w, h, 0, O.CAM_MUSEUM)
    pt.set_config(render_type=W.NORMAL_NEE); orc.mb_config(type=O.NORMAL_NEE, trig=O.TRIG_SHARED)
    # the first queue (4 spp, sampling_strategy.rs:197-203), then the whole first adaptive round: its size is known
    # only after the error map, so render it in two calls and cut the second adaptive round by the budget
    assert pt.render_adaptive(4 * n) == orc.mb_render_adaptive(4 * n, threads=THREADS) == 4 * n
    assert pt.render_adaptive(1) == orc.mb_render_adaptive(1, threads=THREADS) == 1          # opens round 1
    spp1 = pt.round_spp(); ospp1 = orc.mb_round_spp()
    assert np.array_equal(spp1.ravel(), ospp1) and spp1.min() >= 1 and spp1.max() <= 33
    round1 = int(spp1.sum(dtype=np.uint64))
    budget = (round1 - 1) + n // 2                                                            # rest of round 1 + half a million pixels' worth of round 2
    assert pt.render_adaptive(budget) == orc.mb_render_adaptive(budget, threads=THREADS) == budget
    spp2 = pt.round_spp(); ospp2 = orc.mb_round_spp()
    assert np.array_equal(spp2.ravel(), ospp2) and not np.array_equal(spp1, spp2)             # a second error map was evaluated
    rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
    assert np.array_equal(cnt, ocnt) and int(cnt.sum(dtype=np.uint64)) == 4 * n + 1 + budget
    assert np.array_equal(bits(rgb), bits(orgb))
    mse, st = pt.error_map(); omse, ost = orc.mb_error_map(w, h)
    assert np.array_equal(bits(mse), bits(omse)) and np.array_equal(bits(st), bits(ost))
    assert np.array_equal(pt.results(0), orc.results(0)) and np.array_equal(pt.results(1), orc.results(1))
