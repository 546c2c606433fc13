"""The scheduling options of the persistent kernel change the ORDER in which slots are traced, never a result: graded end zones
of the slot queue (WPT_MEGA_ZONES, per-sample colours + segment sums formed by k_combine_segments) and the slot order by
primary-hit class (WPT_TILE_ORDER). Both are read once per process, so every case runs in a child process that renders on
the GPU with the option set and compares the f32 accumulators bit for bit with the oracle's mode B."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CHILD = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
import oracle_lib as O
import wasm_pathtracer_b200 as W
bits = lambda a: np.ascontiguousarray(a).view(np.uint32)
scene, cam, sub, rtype, bvh, spp, w, h = %(scene)d, %(cam)s, %(sub)r, %(rtype)d, %(bvh)d, %(spp)r, %(w)d, %(h)d
pt = W.PathTracer(w, h, scene, *cam, device=0); orc = O.Oracle(w, h, scene, cam)
if sub:
    v = W.parse_obj(open(os.path.join(%(root)r, "assets", "_gen", "standin_%%d.obj" %% sub)).read(), True)
    pt.store_mesh(1, v); orc.load_mesh(1, v)
pt.set_config(render_type=rtype, bvh_kind=bvh, photon_target=20000); orc.mb_config(type=rtype, photon_target=20000)
if bvh == 4: orc.rebuild_bvh(True)
if rtype == 2: pt.build_photons(); orc.mb_build_photons(threads=8)
for n in spp if isinstance(spp, tuple) else (spp,):
    pt.render_exact(n); orc.mb_render_exact(n, threads=8)
a, c = pt.accum(); oa, oc = orc.accum()
assert np.array_equal(c, oc), "sample counts differ"
assert np.array_equal(bits(a), bits(oa)), "accumulator bits differ: %%d pixels" %% int((bits(a) != bits(oa)).any(axis=-1).sum())
st, ost = pt.stats(), orc.stats(0)
assert (st["rays"], st["paths"], st["node_visits"]) == (ost["rays"], ost["paths"], ost["node_visits"]), (st, ost)
print("SAME", st["rays"])
"""

CAM_BUNNY = (-0.9, 5.4, 0.4, 0.58, 0.0)
CAM_MUSEUM = (0.0, 16.34, -23.76, 0.54, 0.0)


def run_child(env_extra, **kw):
    env = dict(os.environ); env.update(env_extra)
    kw["root"] = ROOT
    r = subprocess.run([sys.executable, "-c", CHILD % kw], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SAME" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


@pytest.mark.parametrize("zones", ["30:4,20:3,10:1", "50:1", "20:5,40:2"])
def test_end_zones_leave_every_bit_alone_bunny(gpu_ok, meshes, zones):
    # 20 spp = segments of 8 + 8 + 4; zone slot lengths that divide the segment, that do not, and single samples;
    # the 168 x 100 frame has ragged strips behind the full 8x4 tiles, which lie inside the zones
    # (WPT_NO_SIMPLE: the triangles / planes kernel variants are compiled without the zone code, the generic variant has it)
    run_child({"WPT_MEGA_ZONES": zones, "WPT_TILE_ORDER": "1", "WPT_NO_SIMPLE": "1"}, scene=2, cam=CAM_BUNNY, sub=3, rtype=1, bvh=2, spp=20, w=168, h=100)


def test_end_zones_with_photon_nee_bvh4_and_two_calls(gpu_ok, meshes):
    # two render_exact calls: the second one's samples continue the pixel's sample indices (zone slots read the count too)
    run_child({"WPT_MEGA_ZONES": "25:4,15:2,10:1", "WPT_NO_SIMPLE": "1"}, scene=2, cam=CAM_BUNNY, sub=3, rtype=2, bvh=4, spp=(11, 9), w=96, h=64)


def test_default_zones_and_tile_order_on_the_museum(gpu_ok):
    # scenes with tori have end zones by default (context.cpp); plus the slot order by primary-hit class
    run_child({"WPT_TILE_ORDER": "1"}, scene=0, cam=CAM_MUSEUM, sub=None, rtype=1, bvh=2, spp=12, w=128, h=72)
    run_child({"WPT_MEGA_ZONES": "", "WPT_TILE_ORDER": "0"}, scene=0, cam=CAM_MUSEUM, sub=None, rtype=1, bvh=2, spp=12, w=128, h=72)
