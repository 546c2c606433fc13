"""ctypes binding of oracle/liboracle.so — the CPU restatement of the reference.

TEST INFRASTRUCTURE ONLY (see oracle/ref_core.h): imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.environ.get("WPT_ORACLE_LIBRARY") or os.path.join(ORACLE_DIR, "liboracle.so")   # override: a sanitizer build (profiles/r2_checks.md)

SCENE_MUSEUM, SCENE_BUNNY = 0, 2
SCENE_EXT_WHITTED = 256                         # extension scene
CAM_WHITTED = (0.0, 2.5, -6.0, 0.35, 0.0)
NO_NEE, NORMAL_NEE, PNEE = 0, 1, 2
TRIG_LIBM, TRIG_SHARED = 0, 1
CAM_MUSEUM = (0.0, 16.34, -23.76, 0.54, 0.0)    # src_ts/client/index.ts:156
CAM_BUNNY = (-0.9, 5.4, 0.4, 0.58, 0.0)         # src_ts/client/index.ts:158


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".h", ".cpp", "Makefile"))]
    if os.environ.get("WPT_ORACLE_LIBRARY"):
        return LIB_PATH
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"], stdout=sys.stderr)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint32] * 3 + [C.c_float] * 5 + [C.c_int]
        L.orc_mesh_vertices.restype = C.POINTER(C.c_float)
        L.orc_results.restype = C.POINTER(C.c_uint8)
        L.orc_mb_render_adaptive.restype = C.c_int64
        L.orc_mb_photon_count.restype = C.c_uint64
        L.orc_mb_photon_shots.restype = C.c_uint64
        L.orc_mb_photon_tree.restype = C.c_uint64
        L.orc_parse_obj.restype = C.c_int64
        L.orc_stream_seed.restype = C.c_uint32
        L.orc_museum_colors.restype = C.c_uint32
        L.orc_rng_range.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class OracleError(RuntimeError):
    pass


class Oracle:
    """One reference session (wasm_interface.rs state) with mode-A and mode-B drivers."""

    def __init__(self, width, height, scene_id, camera, bvh4=False):
        self.L = lib()
        self.W, self.H = width, height
        self.h = self.L.orc_create(width, height, scene_id, *[C.c_float(c) for c in camera], int(bvh4))
        if not self.h:
            raise OracleError(self.L.orc_last_error().decode())
        self.h = C.c_void_p(self.h)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc is None or rc < 0:
            raise OracleError(self.L.orc_last_error().decode())
        return rc

    # ---- wasm_interface mirror
    def update_scene(self, sid): self._chk(self.L.orc_update_scene(self.h, sid))
    def update_settings(self, lt, rt, la, ra, dbg): self._chk(self.L.orc_update_settings(self.h, lt, rt, la, ra, dbg))
    def update_viewport(self, w, h): self._chk(self.L.orc_update_viewport(self.h, w, h)); self.W, self.H = w, h
    def update_camera(self, x, y, z, rx, ry): self._chk(self.L.orc_update_camera(self.h, *[C.c_float(v) for v in (x, y, z, rx, ry)]))
    def compute(self, n): self._chk(self.L.orc_compute(self.h, C.c_uint64(n)))
    def reset(self): self._chk(self.L.orc_reset(self.h))
    def rebuild_bvh(self, bvh4): self._chk(self.L.orc_rebuild_bvh(self.h, int(bvh4)))
    def set_trig_a(self, trig): self.L.orc_set_trig_a(self.h, trig)

    def store_texture(self, tex_id, rgb):
        """rgb: uint8 array (h, w, 3) (worker.ts:182-190)."""
        t = np.ascontiguousarray(rgb, dtype=np.uint8)
        self._chk(self.L.orc_store_texture(self.h, tex_id, t.shape[1], t.shape[0], t.ctypes.data_as(C.POINTER(C.c_uint8))))

    def load_mesh(self, mesh_id, verts):
        """verts: float32 array (num_vertices, 3), 3 vertices per triangle (worker.ts:171-179)."""
        v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1)
        nv = v.size // 3
        self._chk(self.L.orc_allocate_mesh(self.h, mesh_id, nv))
        ptr = self.L.orc_mesh_vertices(self.h, mesh_id)
        C.memmove(ptr, v.ctypes.data, v.nbytes)
        return self._chk(self.L.orc_notify_mesh_loaded(self.h, mesh_id)) == 1

    def results(self, show_sampling=0):
        ptr = self.L.orc_results(self.h, show_sampling)
        return np.ctypeslib.as_array(ptr, shape=(self.H, self.W, 4)).copy()

    # ---- mode B
    def mb_config(self, type=NORMAL_NEE, light_debug=False, trig=TRIG_SHARED, seed=0xBABABEBE, photon_target=300000, region=(0, 0, 0, 0)):
        self._chk(self.L.orc_mb_config(self.h, type, int(light_debug), trig, C.c_uint32(seed), C.c_uint64(photon_target), *region))

    def mb_build_photons(self, threads=1): self._chk(self.L.orc_mb_build_photons(self.h, threads))
    def mb_render_exact(self, spp, threads=1): self._chk(self.L.orc_mb_render_exact(self.h, spp, threads))
    def mb_render_adaptive(self, budget, threads=1): return self._chk(self.L.orc_mb_render_adaptive(self.h, C.c_uint64(budget), threads))

    def mb_render_random(self, ticks, threads=1): self._chk(self.L.orc_mb_render_random(self.h, C.c_uint64(ticks), threads))

    def mb_primary_probe(self):
        n = self.W * self.H
        ids = np.empty(n, np.int32); vis = np.empty(n, np.uint32); dist = np.empty(n, np.float32)
        self._chk(self.L.orc_mb_primary_probe(self.h, _p(ids, C.c_int32), _p(vis, C.c_uint32), _p(dist, C.c_float)))
        return ids.reshape(self.H, self.W), vis.reshape(self.H, self.W), dist.reshape(self.H, self.W)

    def mb_round_spp(self):
        n = self.L.orc_mb_round_spp(self.h, None, C.c_uint64(0))
        out = np.empty(n, np.uint32)
        self.L.orc_mb_round_spp(self.h, _p(out, C.c_uint32), C.c_uint64(n))
        return out

    def mb_error_map(self, rw, rh):
        mse = np.empty(rw * rh, np.float32); st = np.empty(3, np.float32)
        self._chk(self.L.orc_mb_error_map(self.h, _p(mse, C.c_float), _p(st, C.c_float)))
        return mse.reshape(rh, rw), st

    # ---- read-backs
    def accum(self):
        n = self.W * self.H
        rgb = np.empty(n * 3, np.float32); cnt = np.empty(n, np.uint32)
        self.L.orc_accum(self.h, _p(rgb, C.c_float), _p(cnt, C.c_uint32))
        return rgb.reshape(self.H, self.W, 3), cnt.reshape(self.H, self.W)

    def stats(self, which=0):
        out = np.zeros(8, np.uint64)
        self.L.orc_stats(self.h, which, _p(out, C.c_uint64))
        return dict(rays=int(out[0]), paths=int(out[1]), node_visits=int(out[2]), photons_shot=int(out[3]),
                    photons_stored=int(out[4]), prim_tests=int(out[5]))

    def scene_info(self):
        out = np.zeros(8, np.uint64)
        self.L.orc_scene_info(self.h, _p(out, C.c_uint64))
        k = ["num_shapes", "num_inf", "num_lights", "bvh2_nodes", "bvh2_depth", "bvh4_nodes", "bvh_kind", "bvh4_depth"]
        return {a: int(b) for a, b in zip(k, out)}

    def bvh2(self):
        n = self.scene_info()["bvh2_nodes"]
        b = np.empty(n * 6, np.float32); lf = np.empty(n, np.uint32); cnt = np.empty(n, np.uint32)
        self.L.orc_bvh2(self.h, _p(b, C.c_float), _p(lf, C.c_uint32), _p(cnt, C.c_uint32))
        return b.reshape(n, 6), lf, cnt

    def bvh4(self):
        n = self.scene_info()["bvh4_nodes"]
        b = np.empty(n * 24, np.float32); ch = np.empty(n * 4, np.int32); nc = np.empty(n, np.uint32)
        self.L.orc_bvh4(self.h, _p(b, C.c_float), _p(ch, C.c_int32), _p(nc, C.c_uint32))
        return b.reshape(n, 4, 6), ch.reshape(n, 4), nc

    def shape_order(self):
        n = self.scene_info()["num_shapes"]
        src = np.empty(n, np.int32); typ = np.empty(n, np.int32)
        self.L.orc_shape_order(self.h, _p(src, C.c_int32), _p(typ, C.c_int32))
        return src, typ

    def lights(self):
        n = self.scene_info()["num_lights"]
        out = np.empty(n, np.uint32)
        self.L.orc_lights(self.h, _p(out, C.c_uint32))
        return out

    def verify_bvh(self): return self.L.orc_verify_bvh(self.h) == 1

    def trace_rays(self, origins, dirs, want_normals=True):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3); d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = len(o)
        ids = np.empty(n, np.int32); dist = np.empty(n, np.float32); vis = np.empty(n, np.uint32)
        nrm = np.empty((n, 3), np.float32) if want_normals else None
        self._chk(self.L.orc_trace_rays(self.h, _p(o, C.c_float), _p(d, C.c_float), C.c_uint64(n), _p(ids, C.c_int32), _p(dist, C.c_float), _p(vis, C.c_uint32), _p(nrm, C.c_float)))
        return ids, dist, vis, nrm

    # ---- photons
    def mb_photons(self):
        n = int(self.L.orc_mb_photon_count(self.h))
        light = np.empty(n, np.uint32); loc = np.empty(n * 3, np.float32); w = np.empty(n, np.float32)
        self.L.orc_mb_photon_list(self.h, _p(light, C.c_uint32), _p(loc, C.c_float), _p(w, C.c_float))
        return light, loc.reshape(n, 3), w, int(self.L.orc_mb_photon_shots(self.h))

    def mb_photon_tree(self):
        n = int(self.L.orc_mb_photon_tree(self.h, None, None, None))
        L_ = self.scene_info()["num_lights"]
        meta = np.empty(n * 3, np.uint32); cum = np.empty(n * L_, np.float32); bins = np.empty(n * L_, np.float32)
        self.L.orc_mb_photon_tree(self.h, _p(meta, C.c_uint32), _p(cum, C.c_float), _p(bins, C.c_float))
        return meta.reshape(n, 3), cum.reshape(n, L_), bins.reshape(n, L_)

    def mb_photon_sample(self, pts, seeds):
        p = np.ascontiguousarray(pts, np.float32).reshape(-1, 3); s = np.ascontiguousarray(seeds, np.uint32)
        n = len(p)
        light = np.empty(n, np.uint32); pdf = np.empty(n, np.float32)
        self._chk(self.L.orc_mb_photon_sample(self.h, _p(p, C.c_float), _p(s, C.c_uint32), C.c_uint64(n), _p(light, C.c_uint32), _p(pdf, C.c_float)))
        return light, pdf


# ---- free functions (known-answer helpers)
def rng_u32(seed, n):
    out = np.empty(n, np.uint32); lib().orc_rng_u32(C.c_uint32(seed), n, _p(out, C.c_uint32)); return out


def rng_f32(seed, n):
    out = np.empty(n, np.float32); lib().orc_rng_f32(C.c_uint32(seed), n, _p(out, C.c_float)); return out


def rng_range(seed, n, lo, hi):
    out = np.empty(n, np.uint32); st = lib().orc_rng_range(C.c_uint32(seed), n, lo, hi, _p(out, C.c_uint32)); return out, st


def stream_seed(index, sample, stream, base=0xBABABEBE):
    return int(lib().orc_stream_seed(C.c_uint32(index), C.c_uint32(sample), C.c_uint32(stream), C.c_uint32(base)))


def museum_colors():
    out = np.empty(27, np.int32); st = lib().orc_museum_colors(_p(out, C.c_int32)); return out.reshape(3, 9), int(st)


def shared_sincos(a):
    a = np.ascontiguousarray(a, np.float32); s = np.empty_like(a); c = np.empty_like(a)
    lib().orc_shared_sincos(_p(a, C.c_float), a.size, _p(s, C.c_float), _p(c, C.c_float)); return s, c


def shared_exp_neg(x):
    x = np.ascontiguousarray(x, np.float32); out = np.empty_like(x)
    lib().orc_shared_exp_neg(_p(x, C.c_float), x.size, _p(out, C.c_float)); return out


def hemisphere(seed, n, normal):
    nn = np.ascontiguousarray(normal, np.float32); out = np.empty((n, 3), np.float32)
    lib().orc_hemisphere(C.c_uint32(seed), n, _p(nn, C.c_float), _p(out, C.c_float)); return out


def empirical_pdf(bins, seed, n):
    b = np.ascontiguousarray(bins, np.float32); hist = np.empty(b.size, np.uint32); probs = np.empty(b.size, np.float32)
    lib().orc_empirical_pdf(_p(b, C.c_float), b.size, C.c_uint32(seed), n, _p(hist, C.c_uint32), _p(probs, C.c_float)); return hist, probs


def shared_f64(which, x):
    """which: 0 cos, 1 acos, 2 cbrt — the shared f64 routines of the quartic solver (deviation B11)."""
    x = np.ascontiguousarray(x, np.float64); out = np.empty_like(x)
    lib().orc_shared_f64(which, _p(x, C.c_double), x.size, _p(out, C.c_double)); return out


def quartic(coef):
    c = np.ascontiguousarray(coef, np.float64); out = np.empty(4, np.float64)
    n = lib().orc_quartic(_p(c, C.c_double), _p(out, C.c_double)); return out[:n]


def parse_obj(text, client_scale=True):
    b = text.encode() if isinstance(text, str) else text
    n = lib().orc_parse_obj(b, C.c_uint64(len(b)), int(client_scale), None, C.c_uint64(0))
    if n < 0:
        raise OracleError(lib().orc_last_error().decode())
    out = np.empty(n, np.float32)
    lib().orc_parse_obj(b, C.c_uint64(len(b)), int(client_scale), _p(out, C.c_float), C.c_uint64(n))
    return out.reshape(-1, 3)
