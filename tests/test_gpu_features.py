"""GPU parity (pytest -m gpu) of the photon warm-up, the octree light CDF, PNEE, the adaptive and
random sampling strategies and the reference-style compute() driver — all bit-exact vs the
oracle's mode B."""
import numpy as np
import pytest

import oracle_lib as O
import wasm_pathtracer_b200 as W

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def pair(scene, cam, w, h, verts=None, photon_target=20000, rtype=W.PNEE, region=(0, 0, 0, 0)):
    pt = W.PathTracer(w, h, scene, *cam, device=0)
    orc = O.Oracle(w, h, scene, cam)
    if verts is not None:
        pt.store_mesh(1, verts); orc.load_mesh(1, verts)
    pt.set_config(render_type=rtype, photon_target=photon_target, region_x=region[0], region_y=region[1], region_w=region[2], region_h=region[3])
    orc.mb_config(type=rtype, photon_target=photon_target, region=region)
    return pt, orc


@pytest.mark.parametrize("scene,cam,sub,target", [(2, W.CAM_BUNNY, 3, 30000), (0, W.CAM_MUSEUM, None, 40000)])
def test_photon_warmup_and_octree(gpu_ok, meshes, scene, cam, sub, target):
    pt, orc = pair(scene, cam, 32, 32, meshes[sub] if sub else None, target)
    pt.build_photons(); orc.mb_build_photons(threads=4)
    light, loc, w, shots = pt.photons()
    olight, oloc, ow, oshots = orc.mb_photons()
    assert len(light) == target and shots == oshots
    assert np.array_equal(light, olight) and np.array_equal(bits(loc), bits(oloc)) and np.array_equal(bits(w), bits(ow))
    meta, cum, bins = pt.photon_tree()
    ometa, ocum, obins = orc.mb_photon_tree()
    assert np.array_equal(meta, ometa)            # same topology: depth, node/leaf, photons per leaf
    assert meta[:, 1].sum() > 0                    # the tree did split
    assert np.array_equal(bits(bins), bits(obins))  # fixed-point bins: order independent, bit-exact
    assert np.array_equal(bits(cum), bits(ocum))
    leaf_photons = meta[meta[:, 1] == 0, 2]
    assert leaf_photons.sum() == target and leaf_photons.max() <= 1024
    st, ost = pt.stats(), orc.stats(0)
    assert (st["photons_shot"], st["photons_stored"], st["rays"], st["node_visits"]) == (ost["photons_shot"], ost["photons_stored"], ost["rays"], ost["node_visits"])
    # light choice (photon_tree.rs:80-159) on random points, incl. points outside the +-1024 cube
    rng = np.random.default_rng(3)
    pts = rng.uniform(-6, 6, (5000, 3)).astype(np.float32) + np.array([0, 1, 4], np.float32)
    pts[:20] *= 500.0
    seeds = rng.integers(1, 2 ** 32 - 1, 5000, dtype=np.uint32)
    l, p = pt.photon_sample(pts, seeds); ol, op = orc.mb_photon_sample(pts, seeds)
    assert np.array_equal(l, ol) and np.array_equal(bits(p), bits(op))


@pytest.mark.parametrize("scene,cam,sub", [(2, W.CAM_BUNNY, 3), (0, W.CAM_MUSEUM, None)])
def test_pnee_radiance_bit_exact(gpu_ok, meshes, scene, cam, sub):
    pt, orc = pair(scene, cam, 96, 64, meshes[sub] if sub else None, 30000)
    pt.render_exact(2); orc.mb_render_exact(2, threads=4)
    assert np.array_equal(bits(pt.accum()[0]), bits(orc.accum()[0]))
    st, ost = pt.stats(), orc.stats(0)
    assert (st["rays"], st["paths"], st["node_visits"]) == (ost["rays"], ost["paths"], ost["node_visits"])
    # update_camera keeps the photons (tracer.rs:84-88): no second warm-up
    pt.update_camera(*cam); orc.update_camera(*cam)
    pt.render_exact(1); orc.mb_render_exact(1)
    assert pt.stats()["photons_shot"] == 0
    assert np.array_equal(bits(pt.accum()[0]), bits(orc.accum()[0]))


def test_adaptive_rounds_bit_exact(gpu_ok, meshes):
    w, h = 80, 48
    pt, orc = pair(0, W.CAM_MUSEUM, w, h, None, rtype=W.NORMAL_NEE)
    n = w * h
    # budgets: inside the first queue, to its end + into round 1, a whole round and a bit, tiny
    for budget in (n * 2 + 7, n * 3, n * 20, 5):
        used = pt.render_adaptive(budget); oused = orc.mb_render_adaptive(budget, threads=4)
        assert used == oused == budget
        rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
        assert np.array_equal(cnt, ocnt)
        assert np.array_equal(bits(rgb), bits(orgb))
        assert np.array_equal(pt.round_spp().ravel(), orc.mb_round_spp())
    mse, st = pt.error_map(); omse, ost = orc.mb_error_map(w, h)
    assert np.array_equal(bits(mse), bits(omse)) and np.array_equal(bits(st), bits(ost))
    spp = pt.round_spp()
    assert spp.min() >= 1 and spp.max() <= 33          # sampling_strategy.rs:163
    assert np.array_equal(pt.results(1), orc.results(1))   # the sampling-density view (mix_color)
    assert np.array_equal(pt.results(0), orc.results(0))


def test_adaptive_region_and_filter_across_the_boundary(gpu_ok, meshes):
    """Right-half region: the Gaussians read across the region boundary (render_target.rs:132-138, quirk q9)."""
    w, h = 64, 40
    region = (32, 0, 32, 40)
    pt, orc = pair(2, W.CAM_BUNNY, w, h, meshes[3], rtype=W.NORMAL_NEE, region=region)
    # give the left half some samples first so that the filter sees real data there
    pt.set_config(region_x=0, region_w=32); orc.mb_config(type=O.NORMAL_NEE, region=(0, 0, 32, 40))
    pt.render_exact(2); orc.mb_render_exact(2)
    pt.set_config(region_x=32, region_w=32); orc.mb_config(type=O.NORMAL_NEE, region=region)
    for budget in (32 * 40 * 4, 32 * 40 * 9):
        assert pt.render_adaptive(budget) == orc.mb_render_adaptive(budget) == budget
    rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
    assert np.array_equal(cnt, ocnt) and np.array_equal(bits(rgb), bits(orgb))
    assert (cnt[:, :32] == 2).all()


def test_random_strategy_bit_exact(gpu_ok, meshes):
    pt, orc = pair(2, W.CAM_BUNNY, 64, 48, meshes[3], rtype=W.NORMAL_NEE)
    for ticks in (1000, 64 * 48 * 3, 1):
        pt.render_random(ticks); orc.mb_render_random(ticks, threads=4)
        rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
        assert np.array_equal(cnt, ocnt) and np.array_equal(bits(rgb), bits(orgb))
    assert pt.accum()[1].sum() == 1000 + 64 * 48 * 3 + 1


def test_random_strategy_many_samples_per_pixel(gpu_ok, meshes):
    """A random-strategy call that gives pixels far more than 64 samples runs in passes of <= 64 per pixel (8 segments of
    8, contract B10) — the same segments in the same order as the oracle's single loop; all engines agree."""
    w, h = 24, 16
    pt, orc = pair(2, W.CAM_BUNNY, w, h, meshes[3], rtype=W.NORMAL_NEE)
    ticks = w * h * 150
    pt.render_random(ticks); orc.mb_render_random(ticks, threads=4)
    rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
    assert cnt.max() > 128 and np.array_equal(cnt, ocnt)
    assert np.array_equal(bits(rgb), bits(orgb))
    for engine in (1,):
        pt.reset(); pt.set_config(engine=engine); pt.render_random(ticks)
        assert np.array_equal(bits(pt.accum()[0]), bits(rgb)), engine
    pt.close()


def test_compute_reference_defaults(gpu_ok, meshes):
    """compute(n) with the reference's default halves: left NormalNEE + random, right PNEE + adaptive
    (wasm_interface.rs:90-97,374-384), replayed on the oracle's mode-B drivers."""
    w, h, n = 64, 32, 64 * 32 * 12
    pt = W.PathTracer(w, h, 2, *W.CAM_BUNNY, device=0)
    orc = O.Oracle(w, h, 2, O.CAM_BUNNY)
    pt.store_mesh(1, meshes[3]); orc.load_mesh(1, meshes[3])
    pt.set_config(photon_target=20000)
    pt.compute(n)
    orc.mb_config(type=O.NORMAL_NEE, photon_target=20000, region=(0, 0, 32, 32))
    orc.mb_render_random(n // 2, threads=4)
    orc.mb_config(type=O.PNEE, photon_target=20000, region=(32, 0, 32, 32))
    orc.mb_build_photons(threads=4)
    shots = orc.mb_photons()[3]
    orc.mb_render_adaptive(n - n // 2 - shots // 32, threads=4)
    rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
    assert np.array_equal(cnt, ocnt) and np.array_equal(bits(rgb), bits(orgb))
    assert np.array_equal(pt.results(0), orc.results(0))
    assert cnt[:, :32].sum() == n // 2
    # update_settings: both halves NoNEE + random, light debug off (wasm_interface.rs:173-204)
    pt.update_settings(0, 0, 0, 0, 0)
    assert pt.accum()[1].sum() == 0
    pt.compute(1000)
    assert pt.accum()[1].sum() == 1000


def test_wavefront_engine_matches_persistent_engine(gpu_ok, meshes):
    a = W.PathTracer(128, 72, 2, *W.CAM_BUNNY, device=0); a.store_mesh(1, meshes[4])
    a.set_config(render_type=W.PNEE, photon_target=20000, engine=0)
    a.build_photons(); a.reset()          # keep the warm-up rays out of the comparison
    a.render_exact(3)
    rgb0, _ = a.accum(); st0 = a.stats()
    a.reset(); a.set_config(engine=1); a.render_exact(3)
    rgb1, _ = a.accum(); st1 = a.stats()
    assert np.array_equal(bits(rgb0), bits(rgb1))
    assert (st0["rays"], st0["node_visits"], st0["paths"]) == (st1["rays"], st1["node_visits"], st1["paths"])


def test_photon_warmup_split_over_ranks_is_bit_exact(gpu_ok, meshes):
    """Multi-GPU photon warm-up, emulated with two sessions on one GPU: rank r emits every 2nd shot of each
    batch and the batch's per-shot slots are merged with an integer sum (the NCCL allreduce of dist.attach).
    Both ranks must end up with exactly the single-session photon list and tree."""
    import threading
    import torch
    from wasm_pathtracer_b200.dist import device_tensor

    ref = W.PathTracer(64, 48, 0, *W.CAM_MUSEUM, device=0)
    ref.set_config(render_type=W.PNEE, photon_target=30000)
    ref.build_photons()
    want = ref.photons(); want_tree = ref.photon_tree(); st_ref = ref.stats()

    world = 2
    barrier = threading.Barrier(world)
    staged = [None] * world
    results = [None] * world
    errors = []

    def rank_main(rank):
        try:
            pt = W.PathTracer(64, 48, 0, *W.CAM_MUSEUM, device=0)
            pt.set_config(render_type=W.PNEE, photon_target=30000, rank=rank, world=world)

            def reduce(ptr, n):   # what ncclAllReduce(uint32, sum) does, with the other session on the same GPU
                pt.synchronize()
                staged[rank] = device_tensor(ptr, (n,), torch.int32)
                barrier.wait()
                total = staged[0].clone()
                for r in range(1, world):
                    total += staged[r]
                torch.cuda.synchronize()
                barrier.wait()
                staged[rank].copy_(total)
                torch.cuda.synchronize()
                barrier.wait()

            pt.set_reduce_callback(reduce)
            pt.build_photons()
            results[rank] = (pt.photons(), pt.photon_tree(), pt.stats())
            pt.close()
        except Exception as e:   # pragma: no cover
            errors.append(e)
            barrier.abort()

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in ts: t.start()
    for t in ts: t.join(120)
    assert not errors, errors
    rays = 0
    for (ph, tree, st) in results:
        assert ph[3] == want[3]
        for a, b in zip(ph[:3], want[:3]):
            assert np.array_equal(bits(a) if a.dtype == np.float32 else a, bits(b) if b.dtype == np.float32 else b)
        for a, b in zip(tree, want_tree):
            assert np.array_equal(bits(a) if a.dtype == np.float32 else a, bits(b) if b.dtype == np.float32 else b)
        rays += st["rays"]
    assert rays == st_ref["rays"]      # each shot was traced by exactly one rank


def test_segmented_accumulation_contract_b10(gpu_ok, meshes):
    """render_exact sums a pixel's samples in segments of 8 (each from +0, added in order), so that one pixel's
    samples can run on several lanes: 37 + 5 samples = segments 8 | 8 | 8 | 8 | 5 then 5. The oracle follows the same
    contract; the three engines and a two-rank partition give the same bits."""
    w, h = 80, 45
    pt, orc = pair(2, W.CAM_BUNNY, w, h, meshes[3], rtype=W.NORMAL_NEE)
    pt.render_exact(37); pt.render_exact(5)
    orc.mb_render_exact(37, threads=4); orc.mb_render_exact(5, threads=4)
    rgb, cnt = pt.accum(); orgb, ocnt = orc.accum()
    assert np.array_equal(cnt, ocnt) and int(cnt.min()) == 42
    # more than 64 samples in one call run as several launches of <= 8 segments: still the same segments in the same order
    small, osmall = pair(2, W.CAM_BUNNY, 24, 16, meshes[3], rtype=W.NORMAL_NEE)
    small.render_exact(150); osmall.mb_render_exact(150, threads=4)
    assert np.array_equal(bits(small.accum()[0]), bits(osmall.accum()[0])) and int(small.accum()[1].min()) == 150
    small.close()
    assert np.array_equal(bits(rgb), bits(orgb))
    st, ost = pt.stats(), orc.stats(0)
    assert (st["rays"], st["node_visits"], st["paths"]) == (ost["rays"], ost["node_visits"], ost["paths"])
    for engine in (1,):
        pt.reset(); pt.set_config(engine=engine); pt.render_exact(37); pt.render_exact(5)
        r2, c2 = pt.accum()
        assert np.array_equal(bits(rgb), bits(r2)) and np.array_equal(cnt, c2), engine
    # two ranks, each rendering its 4-row bands: together the same frame
    parts = []
    for rank in range(2):
        q = W.PathTracer(w, h, 2, *W.CAM_BUNNY, device=0); q.store_mesh(1, meshes[3])
        q.set_config(render_type=W.NORMAL_NEE, rank=rank, world=2)
        q.render_exact(37); q.render_exact(5)
        parts.append(q.accum()[0]); q.close()
    from wasm_pathtracer_b200.dist import rows_of_rank
    merged = parts[0].copy(); r1 = rows_of_rank(h, 1, 2); merged[r1] = parts[1][r1]
    assert np.array_equal(bits(merged), bits(rgb))
    pt.close()
