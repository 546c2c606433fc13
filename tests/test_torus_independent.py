"""Independent check of the torus intersection.

src/graphics/primitives/torus.rs:61-126 solves a quartic with the un-vendored `roots 0.0.4` crate; oracle and GPU share
one restatement of that solver, so their agreement proves nothing about it. Here the returned hit distance is checked
against the torus' implicit equation directly, with no polynomial solver involved:
    F(p) = (|p|^2 + R^2 - r^2)^2 - 4 R^2 (p.x^2 + p.z^2),  p = o - c + t d     (torus.rs:71-99)
  (1) residual: the Newton step |F / F'| at the returned t is below 1e-4 (the root is a root);
  (2) bracketing: F sampled in f64 along the ray has no sign change before t (it is the FIRST root >= 1e-4, torus.rs:130-139),
      and a bisection of the bracketing interval lands within 2e-4 of t;
  (3) rays reported as missing the targeted torus do not enter it on the sampled grid.
Rays are aimed at the museum's 27 tori (src/scenes.rs:36: R = 1.3, r = 0.3 at (x, -0.5, z)) from random points above the floor."""
import numpy as np
import pytest

import oracle_lib as O

R_BIG, R_SMALL = 1.3, 0.3
SH_TORUS = 2


def make_rays(n, seed):
    rng = np.random.default_rng(seed)
    col = rng.integers(0, 9, n); row = rng.integers(0, 3, n)
    c = np.stack([-16.0 + 4.0 * col, np.full(n, -0.5), np.array([-7.5, 0.0, 7.5])[row]], 1)
    o = c + np.stack([rng.uniform(-1.9, 1.9, n), rng.uniform(0.4, 2.4, n), rng.uniform(-2.4, 2.4, n)], 1)
    tgt = c + np.stack([rng.uniform(-1.7, 1.7, n), rng.uniform(-0.35, 0.35, n), rng.uniform(-1.7, 1.7, n)], 1)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32), c


def F(o, d, c, t):
    p = (o - c)[:, None, :] + t[..., None] * d[:, None, :]
    s = (p ** 2).sum(-1) + R_BIG ** 2 - R_SMALL ** 2
    return s * s - 4.0 * R_BIG ** 2 * (p[..., 0] ** 2 + p[..., 2] ** 2)


def check(trace, shape_order, n, seed):
    o32, d32, c = make_rays(n, seed)
    ids, dist, vis, _ = trace(o32, d32)
    src, typ = shape_order()
    o, d = o32.astype(np.float64), d32.astype(np.float64)
    hit_torus = (ids >= 0) & (typ[np.maximum(ids, 0)] == SH_TORUS)
    # the torus that was hit: museum shape order is plane, then per (row, column) a torus followed by 4 light triangles
    k = (src[np.maximum(ids, 0)] - 1) // 5
    own = hit_torus & (np.abs(c[:, 0] - (-16.0 + 4.0 * (k % 9))) < 1e-6) & (np.abs(c[:, 2] - np.array([-7.5, 0.0, 7.5])[np.minimum(k // 9, 2)]) < 1e-6)
    assert own.sum() > 0.25 * n                                          # plenty of rays really hit the torus they aim at
    t = dist.astype(np.float64)
    # (1) residual as a Newton step
    i = np.where(own)[0]
    h = 1e-6
    f0 = F(o[i], d[i], c[i], t[i][:, None])[:, 0]
    f1 = F(o[i], d[i], c[i], (t[i] + h)[:, None])[:, 0]
    fp = (f1 - f0) / h
    ok_slope = np.abs(fp) > 1e-3                                         # skip tangential hits: the step is ill-defined there
    step = np.abs(f0[ok_slope] / fp[ok_slope])
    assert ok_slope.mean() > 0.95 and step.max() < 1e-4, (ok_slope.mean(), step.max())
    # (2) first root: no earlier entry on a fine grid, bisection agrees
    m = 2048
    bad = 0; worst = 0.0
    for a in range(0, len(i), 4096):
        j = i[a:a + 4096]
        grid = 1e-4 + (t[j][:, None] - 2e-3 - 1e-4) * np.linspace(0.0, 1.0, m)[None, :]
        f = F(o[j], d[j], c[j], grid)
        bad += int(((f < -1e-9).any(1) & (t[j] > 5e-3)).sum())          # inside the tube before the reported hit
        lo, hi = t[j] - 2e-3, t[j] + 2e-3                               # the surface is crossed inside +-2e-3 of t
        flo = F(o[j], d[j], c[j], lo[:, None])[:, 0]; fhi = F(o[j], d[j], c[j], hi[:, None])[:, 0]
        cross = (flo > 0) & (fhi < 0)
        for _ in range(40):
            mid = 0.5 * (lo + hi)
            fm = F(o[j], d[j], c[j], mid[:, None])[:, 0]
            neg = fm < 0
            hi = np.where(neg, mid, hi); lo = np.where(neg, lo, mid)
        err = np.abs(0.5 * (lo + hi) - t[j])[cross]
        worst = max(worst, float(err.max()) if err.size else 0.0)
    assert bad == 0 and worst < 2e-4, (bad, worst)
    # (3) rays that do not hit their torus (another shape is nearer, or nothing): they do not enter it before that hit
    j = np.where(~own)[0][:20000]
    tmax = np.where(ids[j] >= 0, t[j], 12.0)
    grid = 1e-4 + (np.maximum(tmax - 2e-3, 2e-4)[:, None] - 1e-4) * np.linspace(0.0, 1.0, m)[None, :]
    f = F(o[j], d[j], c[j], grid)
    assert not (f < -1e-6).any(), int((f < -1e-6).any(1).sum())
    return int(own.sum())


def test_oracle_torus_hits_satisfy_the_implicit_equation(built):
    orc = O.Oracle(32, 32, O.SCENE_MUSEUM, O.CAM_MUSEUM)
    assert check(orc.trace_rays, orc.shape_order, 20000, 11) > 5000


@pytest.mark.gpu
def test_gpu_torus_hits_satisfy_the_implicit_equation(gpu_ok):
    import wasm_pathtracer_b200 as W
    pt = W.PathTracer(32, 32, W.SCENE_MUSEUM, *W.CAM_MUSEUM, device=0)
    assert check(pt.trace_rays, pt.shape_order, 100000, 12) > 25000
