"""Our renders against the images the reference ships (SURVEY 8c iii; VERDICT r1 "missing" #5).

/root/reference cannot be read on the GPU box, so tests/golden/make_ref_thumbs.py (run in the build container) reduced
public_html/images/banners/lights.png, banner.png and bunny_high.png to small signatures in tests/golden/ref_thumbs.npz.
The museum's 27 area lights are directly visible, saturated emitters: the left-to-right colour sequence of every row of
light panels in the reference's renders must appear in ours (tests/panel_rows.py) — that pins the scene-local RNG,
`shuffle`, the colour table and light layout of src/scenes.rs:15-68, the camera model of src/tracer.rs:156-201 and the
emissive branch of src/tracer.rs:245-254 against an artefact produced by the reference itself."""
import os

import numpy as np
import pytest

import oracle_lib as O
from panel_rows import contains, panel_rows

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_thumbs.npz")


def ref_rows(key):
    g = np.load(GOLD)[key]
    return [[tuple(int(v) for v in c) for c in row if c[0] >= 0] for row in g]


def check_rows(img):
    ours = panel_rows(img.astype(np.float32) / 255.0)
    assert len(ours) >= 3 and sum(len(r) == 9 for r in ours) >= 2   # the two far rows of nine lights are fully in view
    for key in ("lights_rows", "banner_rows"):
        for row in ref_rows(key):
            assert any(contains(o, row) for o in ours), (key, row, ours)
    # and the three full rows are exactly the colour order of the KAT (tests/test_oracle_kat.py)
    table = [(2, 1, 1), (0, 2, 2), (1, 1, 2), (2, 0, 0), (0, 2, 0), (0, 0, 2), (2, 0, 2), (2, 2, 0), (1, 2, 1)]   # scenes.rs:22-28 x 2.5, clamped
    order, _ = O.museum_colors()
    for z in range(3):                                                # scene rows z = -7.5 (near: its ends are out of view), 0, 7.5
        kat = [table[i] for i in order[z]]
        assert any(len(r) >= 7 and contains(kat, r) for r in ours), (z, kat, ours)


def test_oracle_museum_render_shows_the_reference_light_order(built):
    o = O.Oracle(500, 500, O.SCENE_MUSEUM, O.CAM_MUSEUM)             # the default camera of src_ts/client/index.ts:156
    o.mb_config(type=O.NO_NEE, trig=O.TRIG_SHARED)
    o.mb_render_exact(4, threads=os.cpu_count() or 1)
    check_rows(o.results(0)[..., :3])


@pytest.mark.gpu
def test_gpu_museum_render_shows_the_reference_light_order(gpu_ok):
    import wasm_pathtracer_b200 as W
    pt = W.PathTracer(500, 500, W.SCENE_MUSEUM, *W.CAM_MUSEUM, device=0)
    pt.set_config(render_type=W.NORMAL_NEE)
    pt.render_exact(16)
    check_rows(pt.results(0)[..., :3])


@pytest.mark.gpu
def test_gpu_bunny_render_has_the_hue_of_the_reference_thumbnail(gpu_ok, meshes):
    """bunny_high.png shows the real bunny2.obj (a stripped blob here), so only what does not depend on the
    mesh shape is compared: the lit mesh is the diffuse (1, 0.4, 0.4) of src/wasm_interface.rs:300-308 under a white
    light — red dominates, green and blue are equal — on a dark floor / wall (the stand-in covers less of the frame than
    the bunny, so the brightest blocks hold more of the grey floor: the ratios are bounded, not matched)."""
    import wasm_pathtracer_b200 as W
    ref = np.load(GOLD)["bunny"]
    pt = W.PathTracer(300, 180, W.SCENE_BUNNY, *W.CAM_BUNNY, device=0)
    pt.store_mesh(1, meshes[4])
    pt.set_config(render_type=W.NORMAL_NEE)
    pt.render_exact(64)
    img = pt.results(0)[..., :3].astype(np.float32) / 255.0
    ours = img[:180, :300].reshape(18, 10, 30, 10, 3).mean((1, 3))

    def lit_hue(g):
        sel = g[..., 0] > np.percentile(g[..., 0], 85)                # the brightest (mesh) blocks
        c = g[sel].mean(0)
        return c[1] / c[0], c[2] / c[0]

    (g1, b1), (g2, b2) = lit_hue(ours), lit_hue(ref)
    assert g1 < 0.75 and g2 < 0.75 and abs(g1 - b1) < 0.05 and abs(g2 - b2) < 0.05 and abs(g1 - g2) < 0.3, ((g1, b1), (g2, b2))
    assert np.median(ours.max(-1)) < 0.35 and np.median(ref.max(-1)) < 0.35   # mostly dark surroundings in both
