"""Radiance gate against the REFERENCE'S OWN semantics (DESIGN.md 3 "Radiance tolerance").

Every other GPU test compares the CUDA path with the oracle's mode B (per-path streams, shared trig, segment sums:
deviations B1, B2, B10) bit for bit. This file bounds those deviations against mode A — the oracle driver that follows
the reference literally: ONE xorshift32 stream shared by pixel choice, jitter, BSDF, light choice and Russian roulette
(src/rng.rs:10-12, src/wasm_interface.rs:87), libm sin / cos (src/graphics/material.rs:103-105), every sample added to
the accumulator as it arrives (src/render_target.rs:55-58), pixels drawn by RandomSamplingStrategy
(src/graphics/sampling_strategy.rs:56-59).

The two renderers cannot share random numbers, so at equal spp they differ by Monte-Carlo noise; what has
to hold is that they estimate the SAME image: the difference falls like 1 / sqrt(spp) without a floor, the frame means
agree, and at 1024 spp the images are within the tolerances stated here (and in DESIGN.md):

  bunny  128x72, NormalNEE, clamped means: whole-image RMSE <= 0.025 @ 256 spp and <= 0.0125 @ 1024 spp;
         per-pixel relative difference |a - b| / max(a, b, 0.05): 99 % of the pixels <= 0.25, all <= 0.45 @ 1024 spp;
         RMSE(64 spp) / RMSE(1024 spp) in [3, 5] (4 = pure 1 / sqrt(N)); frame means within 0.5 %.
  museum 128x72, NormalNEE (108 small lights: heavy-tailed noise), clamped means averaged over 4x4 pixel blocks:
         RMSE <= 0.03 @ 256 spp and <= 0.016 @ 1024 spp; median per-pixel relative difference <= 0.04 @ 1024 spp;
         block RMSE(64) / RMSE(1024) >= 2.5; frame means within 2 %.
(Measured with the oracle's mode B, which the GPU equals bit for bit: bunny 0.0359 / 0.0182 / 0.0092, museum blocks
0.0359 / 0.0214 / 0.0109 at 64 / 256 / 1024 spp.)
"""
import numpy as np
import pytest

import oracle_lib as O
import wasm_pathtracer_b200 as W

pytestmark = pytest.mark.gpu
WID, HEI = 128, 72


def mode_a(scene, cam, verts, spp, rtype=O.NORMAL_NEE):
    """The reference's compute(): both halves NormalNEE + random strategy, one shared stream, libm trig."""
    o = O.Oracle(WID, HEI, scene, cam)
    if verts is not None:
        o.load_mesh(1, verts)
    o.update_settings(rtype, rtype, 0, 0, 0)
    o.set_trig_a(O.TRIG_LIBM)
    o.compute(WID * HEI * spp)
    rgb, cnt = o.accum()
    o.close()
    assert cnt.min() > 0
    return rgb / cnt[..., None]


def gpu(scene, cam, verts, spp, rtype=W.NORMAL_NEE):
    pt = W.PathTracer(WID, HEI, scene, *cam, device=0)
    if verts is not None:
        pt.store_mesh(1, verts)
    pt.set_config(render_type=rtype)
    pt.render_exact(spp)
    rgb, cnt = pt.accum()
    pt.close()
    assert (cnt == spp).all()
    return rgb / cnt[..., None]


def box4(x):
    return x[: HEI // 4 * 4, : WID // 4 * 4].reshape(HEI // 4, 4, WID // 4, 4, 3).mean((1, 3))


def rmse(a, b):
    return float(np.sqrt(((a - b) ** 2).mean()))


def test_bunny_converges_to_the_reference_semantics(gpu_ok, meshes):
    r = {}
    for spp in (64, 256, 1024):
        a = mode_a(2, O.CAM_BUNNY, meshes[4], spp)
        g = gpu(2, W.CAM_BUNNY, meshes[4], spp)
        ac, gc = np.clip(a, 0, 1), np.clip(g, 0, 1)   # RenderTarget::read_clamped, render_target.rs:74-77
        r[spp] = rmse(ac, gc)
        if spp == 1024:
            rel = np.abs(ac - gc).max(-1) / np.maximum(np.maximum(ac, gc).max(-1), 0.05)
            assert np.percentile(rel, 99) <= 0.25 and rel.max() <= 0.45, (np.percentile(rel, 99), rel.max())
            assert abs(a.mean() - g.mean()) <= 0.005 * a.mean(), (a.mean(), g.mean())
    assert r[256] <= 0.025 and r[1024] <= 0.0125, r
    assert 3.0 <= r[64] / r[1024] <= 5.0, r


def test_museum_converges_to_the_reference_semantics(gpu_ok):
    r = {}
    for spp in (64, 256, 1024):
        a = mode_a(0, O.CAM_MUSEUM, None, spp)
        g = gpu(0, W.CAM_MUSEUM, None, spp)
        ac, gc = np.clip(a, 0, 1), np.clip(g, 0, 1)
        r[spp] = rmse(box4(ac), box4(gc))
        if spp == 1024:
            rel = np.abs(ac - gc).max(-1) / np.maximum(np.maximum(ac, gc).max(-1), 0.05)
            assert np.median(rel) <= 0.04, np.median(rel)
            assert abs(a.mean() - g.mean()) <= 0.02 * a.mean(), (a.mean(), g.mean())
    assert r[256] <= 0.03 and r[1024] <= 0.016, r
    assert r[64] / r[1024] >= 2.5, r


def test_pnee_estimates_the_same_image_as_the_reference_semantics(gpu_ok, meshes):
    """PNEE changes the light choice, not the estimate (tracer.rs:270-278): GPU PNEE (mode-B photon tree, fixed-point
    bins — deviation B3) against the reference-order mode A with photon-based NEE on both halves."""
    a = mode_a(2, O.CAM_BUNNY, meshes[4], 256, O.PNEE)
    pt = W.PathTracer(WID, HEI, 2, *W.CAM_BUNNY, device=0)
    pt.store_mesh(1, meshes[4])
    pt.set_config(render_type=W.PNEE)
    pt.build_photons()
    pt.render_exact(256)
    rgb, cnt = pt.accum()
    g = rgb / cnt[..., None]
    assert rmse(np.clip(a, 0, 1), np.clip(g, 0, 1)) <= 0.03
    assert abs(a.mean() - g.mean()) <= 0.01 * a.mean(), (a.mean(), g.mean())
