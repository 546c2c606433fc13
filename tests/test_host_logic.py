"""CPU tests of the product's host side (no GPU): the C ABI loads and exports every declared
symbol; the in-place BVH2 builder, the bottom-up BVH4 collapse, the scene setup and the OBJ
loader agree bit for bit with the oracle's restatement of bvh.rs / bvh4.rs / scenes.rs /
obj_parser.ts."""
import os
import re

import numpy as np
import pytest

import oracle_lib as O
import wasm_pathtracer_b200 as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_abi_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "wpt.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(wpt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) > 50
    L = W.load_library()
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    assert set(L._wpt_symbols) == set(declared)   # the ctypes layer binds exactly the header


def test_no_cpu_fallback(built):
    """Host-only sessions can inspect scenes; anything that would compute fails loudly."""
    pt = W.PathTracer(16, 16, W.SCENE_MUSEUM, *W.CAM_MUSEUM, device=W.DEVICE_NONE)
    for call in (lambda: pt.render_exact(1), lambda: pt.results(0), lambda: pt.primary_probe(), lambda: pt.compute(10), lambda: pt.accum()):
        with pytest.raises(W.WptError, match="no CUDA device|not implemented"):
            call()


def test_error_behaviour_matches_reference_panics(built):
    with pytest.raises(W.WptError, match="Invalid scene"):          # wasm_interface.rs:396
        W.PathTracer(16, 16, 1, *W.CAM_BUNNY, device=W.DEVICE_NONE)
    pt = W.PathTracer(16, 16, W.SCENE_BUNNY, *W.CAM_BUNNY, device=W.DEVICE_NONE)
    with pytest.raises(W.WptError, match="Invalid RenderType"):     # wasm_interface.rs:212
        pt.update_settings(0, 7, 0, 0, 0)
    with pytest.raises(W.WptError, match="Mesh not allocated"):     # wasm_interface.rs:281
        pt.mesh_vertices(5, 3)
    with pytest.raises(W.WptError, match="Invalid scene"):
        pt.update_scene(3)
    assert pt.notify_texture_loaded(0) is False                      # wasm_interface.rs:357-366
    tex = pt.allocate_texture(0, 16, 16)
    assert tex.shape == (16, 16, 3)
    L = W.load_library()
    assert L.wpt_global_ctx() is None
    L.wpt_compute(10)                                                # "init not called" (wasm_interface.rs:381)
    assert b"init not called" in L.wpt_last_error()


@pytest.mark.parametrize("scene,cam,sub", [(0, W.CAM_MUSEUM, None), (2, W.CAM_BUNNY, None), (2, W.CAM_BUNNY, 3), (2, W.CAM_BUNNY, 4)])
def test_bvh2_builder_parity(meshes, scene, cam, sub):
    pt = W.PathTracer(16, 16, scene, *cam, device=W.DEVICE_NONE)
    orc = O.Oracle(16, 16, scene, cam)
    if sub:
        assert pt.store_mesh(1, meshes[sub]) is True
        assert orc.load_mesh(1, meshes[sub]) is True
    assert pt.scene_info() == orc.scene_info()
    for a, b in zip(pt.bvh2(), orc.bvh2()):
        assert np.array_equal(bits(a), bits(b))
    for a, b in zip(pt.shape_order(), orc.shape_order()):
        assert np.array_equal(a, b)
    assert np.array_equal(pt.lights(), orc.lights())
    assert orc.verify_bvh()


@pytest.mark.parametrize("sub", [3, 4])
def test_bvh4_collapse_parity(meshes, sub):
    pt = W.PathTracer(16, 16, 2, *W.CAM_BUNNY, device=W.DEVICE_NONE)
    orc = O.Oracle(16, 16, 2, O.CAM_BUNNY)
    pt.store_mesh(1, meshes[sub]); orc.load_mesh(1, meshes[sub])
    pt.set_config(bvh_kind=4); orc.rebuild_bvh(True)
    assert pt.scene_info() == orc.scene_info()
    ab, ac, an = pt.bvh4()
    bb, bc, bn = orc.bvh4()
    assert np.array_equal(an, bn) and np.array_equal(ac, bc)
    used = np.arange(4)[None, :] < an[:, None]
    assert np.array_equal(bits(ab[used]), bits(bb[used]))
    assert orc.verify_bvh()
    assert (an >= 2).all() and (an <= 4).all()


def test_bvh4_unencodable_scenes_fail_like_the_reference(built):
    pt = W.PathTracer(16, 16, 0, *W.CAM_MUSEUM, device=W.DEVICE_NONE)
    with pytest.raises(W.WptError, match="more than 15 shapes"):     # bvh4.rs:22,135 (finding F7)
        pt.set_config(bvh_kind=4)
    assert pt.get_config().bvh_kind == 2                              # the failed call changed nothing
    pt = W.PathTracer(16, 16, 2, *W.CAM_BUNNY, device=W.DEVICE_NONE)   # mesh-less bunny: the BVH2 root is a leaf
    with pytest.raises(W.WptError, match="root is a leaf"):          # bvh4.rs:67
        pt.set_config(bvh_kind=4)


def test_obj_parser_parity_and_edge_cases(built):
    text = open(os.path.join(ROOT, "assets", "_gen", "standin_3.obj")).read()
    for scale in (True, False):
        assert np.array_equal(bits(W.parse_obj(text, scale)), bits(O.parse_obj(text, scale)))
    v = W.parse_obj(text, True)
    assert v.shape == (1280 * 3, 3)
    raw = W.parse_obj(text, False)
    assert np.array_equal(v, raw * np.array([8, 8, -8], np.float32))   # index.ts:216-220
    # dialect: `f a/b/c`, `f a//c`, comments, unknown records, CRLF, empty input
    t2 = "# c\nvn 0 0 1\nv 0 0 0\nv 1 0 0\r\nv 0 1 0\nusemtl x\nf 1/1/1 2//1 3\n"
    a, b = W.parse_obj(t2, False), O.parse_obj(t2, False)
    assert np.array_equal(a, b) and a.tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0]]
    assert W.parse_obj("", True).shape == (0, 3) and O.parse_obj("", True).shape == (0, 3)
    for bad in ("v 0 0 0\nf 1 1 1 1\n", "f 1 2\n"):                    # obj_parser.ts:27-29
        with pytest.raises(W.WptError, match="Non-triangular"):
            W.parse_obj(bad)
        with pytest.raises(O.OracleError, match="Non-triangular"):
            O.parse_obj(bad)
    # out-of-range / missing indices become NaN like a Float32Array store of `undefined`
    a, b = W.parse_obj("v 1 2 3\nf 1 2 x\n", False), O.parse_obj("v 1 2 3\nf 1 2 x\n", False)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and a[0].tolist() == [1, 2, 3] and np.isnan(a[1:]).all()


def test_load_obj_file_path(built):
    pt = W.PathTracer(16, 16, 2, *W.CAM_BUNNY, device=W.DEVICE_NONE)
    n = pt.load_obj(1, os.path.join(ROOT, "assets", "_gen", "standin_3.obj"))
    assert n == 1280 * 3 and pt.scene_info()["num_shapes"] == 1280 + 4
    with pytest.raises(W.WptError, match="cannot open"):
        pt.load_obj(1, "/nonexistent.obj")


def test_mesh_upload_three_step_protocol(meshes):
    """allocate_mesh / mesh_vertices / notify_mesh_loaded as the worker drives them (worker.ts:171-179)."""
    pt = W.PathTracer(16, 16, W.SCENE_MUSEUM, *W.CAM_MUSEUM, device=W.DEVICE_NONE)
    v = meshes[3]
    pt.allocate_mesh(1, len(v))
    pt.mesh_vertices(1, len(v))[:] = v
    assert pt.notify_mesh_loaded(1) is False        # scene 0 does not use mesh 1 (wasm_interface.rs:316-319)
    assert pt.scene_info()["num_shapes"] == 146
    pt.update_scene(W.SCENE_BUNNY)                  # ... but the triangles are kept for scene 2
    assert pt.scene_info()["num_shapes"] == 1280 + 4
    assert pt.notify_mesh_loaded(1) is True         # already triangled: still reports the scene uses it


def test_stand_in_mesh_is_deterministic(built):
    import hashlib
    from assets.make_standin_mesh import standin
    p, f = standin(3)
    assert p.shape == (642, 3) and f.shape == (1280, 3)
    text = open(os.path.join(ROOT, "assets", "_gen", "standin_3.obj")).read()
    assert hashlib.sha256(text.encode()).hexdigest() == open(os.path.join(ROOT, "tests", "golden", "standin_3.sha256")).read().strip()


def test_predicated_sort_network_is_the_branchy_one():
    """device_core.cuh keeps two forms of `sort_small` (scene.rs:346-388): the reference's branches and, behind -DWPT_SORT_PREDICATED,
    the same compare-and-swap network with predicated swaps. Both written out here in Python, all inputs over a value set with
    ties, infinities and NaN, n = 0..4: same permutation, so the macro changes no visit order."""
    import itertools
    import math

    def branchy(d, n):
        ids, d = [0, 1, 2, 3], list(d)

        def sw(i, j):
            ids[i], ids[j] = ids[j], ids[i]; d[i], d[j] = d[j], d[i]
        if n == 2:
            if d[1] < d[0]: sw(0, 1)
        elif n == 3:
            if d[1] < d[0]: sw(0, 1)
            if d[2] < d[1]: sw(1, 2)
            if d[1] < d[0]: sw(0, 1)
        elif n == 4:
            if d[1] < d[0]: sw(0, 1)
            if d[3] < d[2]: sw(2, 3)
            if d[0] < d[2]:
                if d[2] < d[1]:
                    sw(1, 2)
                    if d[3] < d[2]: sw(2, 3)
            else:
                sw(0, 2); sw(1, 2)
                if d[3] < d[1]:
                    sw(1, 3); sw(2, 3)
                elif d[3] < d[2]:
                    sw(2, 3)
        return ids

    def predicated(d, n):
        ids, d = [0, 1, 2, 3], list(d)

        def cs(i, j, c):
            if c:
                ids[i], ids[j] = ids[j], ids[i]; d[i], d[j] = d[j], d[i]
        n3, n4 = n == 3, n == 4
        cs(0, 1, n >= 2 and d[1] < d[0])
        cs(1, 2, n3 and d[2] < d[1])
        cs(0, 1, n3 and d[1] < d[0])
        cs(2, 3, n4 and d[3] < d[2])
        a, b = n4 and d[0] < d[2], n4 and not (d[0] < d[2])
        a1 = a and d[2] < d[1]
        cs(1, 2, a1)
        cs(2, 3, a1 and d[3] < d[2])
        cs(0, 2, b); cs(1, 2, b)
        b1 = b and d[3] < d[1]
        b2 = b and not b1 and d[3] < d[2]
        cs(1, 3, b1); cs(2, 3, b1 or b2)
        return ids

    vals = [0.0, 0.5, 1.0, 2.0, -math.inf, math.inf, math.nan]
    for n in range(5):
        for d in itertools.product(vals, repeat=4):
            assert branchy(d, n) == predicated(d, n), (n, d)
