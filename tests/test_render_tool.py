"""tools/wpt_render: the worker-style progressive driver (src_ts/worker/worker.ts:55-95) written in C++ against
include/wpt.h only — the C-ABI seen from a compiled caller."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wasm_pathtracer_b200 as W
from wasm_pathtracer_b200.build import build_tools


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_driver_builds_and_fails_loudly_without_a_gpu(built):
    exe = build_tools()
    assert os.path.exists(exe)
    if _have_gpu():
        pytest.skip("a CUDA device is visible: the no-device message cannot be provoked")
    r = subprocess.run([exe, "--ticks", "1", "--quiet"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr      # no CPU fallback behind the ABI


def _read_ppm(path):
    raw = open(path, "rb").read()
    assert raw[:2] == b"P6"
    parts = raw.split(b"\n", 3)
    w, h = map(int, parts[1].split())
    return np.frombuffer(parts[3], np.uint8).reshape(h, w, 3)


@pytest.mark.gpu
def test_driver_tick_equals_the_python_host_layer(gpu_ok, meshes, tmp_path):
    exe = build_tools()
    obj = os.path.join(ROOT, "assets", "_gen", "standin_3.obj")
    out = str(tmp_path / "f.ppm")
    # one tick = compute(1000) (worker.ts starts with numRaysPerTick = 1000); then three more with fixed sizes
    r = subprocess.run([exe, "--scene", "2", "--obj", obj, "--size", "64x48", "--samples", "1000", "--quiet", "--out", out],
                       capture_output=True, text=True, check=True)
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["ticks"] == 1 and info["samples"] == 1000
    pt = W.PathTracer(64, 48, W.SCENE_BUNNY, *W.CAM_BUNNY, device=0)
    pt.store_mesh(1, meshes[3])
    pt.update_settings(W.NORMAL_NEE, W.PNEE, 0, 1, 0)
    pt.compute(1000)
    img = pt.results(0).reshape(48, 64, 4)
    assert np.array_equal(_read_ppm(out), img[:, :, :3])
    st = pt.stats()
    assert (info["rays"], info["paths"], info["photons"]) == (st["rays"], st["paths"], st["photons_stored"])
    # PNG output + sampling view + time-boxed loop
    png = str(tmp_path / "f.png")
    r = subprocess.run([exe, "--scene", "0", "--size", "64x48", "--seconds", "0.3", "--sampling-view", "--quiet", "--out", png],
                       capture_output=True, text=True, check=True)
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["ticks"] >= 1 and open(png, "rb").read(8) == b"\x89PNG\r\n\x1a\n"
