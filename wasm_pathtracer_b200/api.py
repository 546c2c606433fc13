"""ctypes host layer over libwpt.so, mirroring the reference's exported functions.

Every method of `PathTracer` named like a `#[wasm_bindgen] pub fn` of
`src/wasm_interface.rs` takes the same arguments with the same meaning; failures raise
`WptError` where the reference panics (see include/wpt.h for the per-function citations).
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
NO_NEE, NORMAL_NEE, PNEE = 0, 1, 2            # wasm_interface.rs:207-214
SCENE_MUSEUM, SCENE_BUNNY = 0, 2              # wasm_interface.rs:389-398
SCENE_EXT_WHITTED = 256                       # extension scene: not a reference id
CAM_WHITTED = (0.0, 2.5, -6.0, 0.35, 0.0)
CAM_MUSEUM = (0.0, 16.34, -23.76, 0.54, 0.0)  # src_ts/client/index.ts:156
CAM_BUNNY = (-0.9, 5.4, 0.4, 0.58, 0.0)       # src_ts/client/index.ts:158
DEVICE_NONE = -2
MESH_BUNNY_HIGH = 1                           # src/scenes.rs:12, src_ts/client/meshes.ts:9


class WptError(RuntimeError):
    pass


class WptConfig(C.Structure):
    _fields_ = [("bvh_kind", C.c_uint32), ("render_type", C.c_uint32), ("light_debug", C.c_uint32), ("base_seed", C.c_uint32),
                ("photon_target", C.c_uint64), ("region_x", C.c_uint32), ("region_y", C.c_uint32), ("region_w", C.c_uint32),
                ("region_h", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32), ("engine", C.c_uint32), ("reserved", C.c_uint32 * 3)]


def library_path():
    # WPT_LIBRARY: A/B builds of the same library during kernel tuning (scripts/); never a different backend
    return os.environ.get("WPT_LIBRARY") or os.path.join(PKG_DIR, "libwpt.so")


_lib = None


def load_library():
    """Load libwpt.so. Fails loudly when it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise WptError("libwpt.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(path)
    vp, u32, u64, f32, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int
    P = C.POINTER
    sig = {
        "wpt_last_error": (C.c_char_p, []),
        "wpt_global_ctx": (vp, []),
        "wpt_default_config": (None, [P(WptConfig)]),
        "wpt_init": (None, [u32, u32, u32, f32, f32, f32, f32, f32]),
        "wpt_results": (P(C.c_uint8), [u32]),
        "wpt_reset": (None, []),
        "wpt_update_scene": (None, [u32]),
        "wpt_update_settings": (None, [u32] * 5),
        "wpt_update_viewport": (None, [u32, u32]),
        "wpt_update_camera": (None, [f32] * 5),
        "wpt_allocate_mesh": (None, [u32, u32]),
        "wpt_mesh_vertices": (P(f32), [u32]),
        "wpt_notify_mesh_loaded": (i32, [u32]),
        "wpt_allocate_texture": (P(C.c_uint8), [u32, u32, u32]),
        "wpt_notify_texture_loaded": (i32, [u32]),
        "wpt_compute": (None, [u64]),
        "wpt_ctx_create": (vp, [i32, u32, u32, u32, f32, f32, f32, f32, f32]),
        "wpt_ctx_destroy": (None, [vp]),
        "wpt_ctx_results": (P(C.c_uint8), [vp, u32]),
        "wpt_ctx_reset": (i32, [vp]),
        "wpt_ctx_update_scene": (i32, [vp, u32]),
        "wpt_ctx_update_settings": (i32, [vp] + [u32] * 5),
        "wpt_ctx_update_viewport": (i32, [vp, u32, u32]),
        "wpt_ctx_update_camera": (i32, [vp] + [f32] * 5),
        "wpt_ctx_allocate_mesh": (i32, [vp, u32, u32]),
        "wpt_ctx_mesh_vertices": (P(f32), [vp, u32]),
        "wpt_ctx_notify_mesh_loaded": (i32, [vp, u32]),
        "wpt_ctx_allocate_texture": (P(C.c_uint8), [vp, u32, u32, u32]),
        "wpt_ctx_notify_texture_loaded": (i32, [vp, u32]),
        "wpt_ctx_compute": (i32, [vp, u64]),
        "wpt_ctx_set_config": (i32, [vp, P(WptConfig)]),
        "wpt_ctx_get_config": (i32, [vp, P(WptConfig)]),
        "wpt_ctx_render_exact": (i32, [vp, u32]),
        "wpt_ctx_render_adaptive": (C.c_int64, [vp, u64]),
        "wpt_ctx_build_photons": (i32, [vp]),
        "wpt_ctx_render_random": (i32, [vp, u64]),
        "wpt_ctx_set_exchange_callback": (i32, [vp, C.c_void_p, vp]),
        "wpt_ctx_set_reduce_callback": (i32, [vp, C.c_void_p, vp]),
        "wpt_ctx_synchronize": (i32, [vp]),
        "wpt_nccl_unique_id": (i32, [P(C.c_uint8)]),
        "wpt_ctx_attach_nccl": (i32, [vp, P(C.c_uint8), u32, u32]),
        "wpt_ctx_attach_nccl_comm": (i32, [vp, vp, u32, u32]),
        "wpt_ctx_detach_nccl": (i32, [vp]),
        "wpt_ctx_gather_frame": (i32, [vp]),
        "wpt_ctx_stats": (i32, [vp, P(u64)]),
        "wpt_ctx_primary_probe": (i32, [vp, P(C.c_int32), P(u32), P(f32)]),
        "wpt_ctx_accum": (i32, [vp, P(f32), P(u32)]),
        "wpt_ctx_trace_rays": (i32, [vp, P(f32), P(f32), u64, P(C.c_int32), P(f32), P(u32), P(f32)]),
        "wpt_ctx_scene_info": (i32, [vp, P(u64)]),
        "wpt_ctx_bvh2": (i32, [vp, P(f32), P(u32), P(u32)]),
        "wpt_ctx_bvh4": (i32, [vp, P(f32), P(C.c_int32), P(u32)]),
        "wpt_ctx_shape_order": (i32, [vp, P(C.c_int32), P(C.c_int32)]),
        "wpt_ctx_lights": (i32, [vp, P(u32)]),
        "wpt_ctx_photon_count": (C.c_int64, [vp, P(u64)]),
        "wpt_ctx_photon_list": (i32, [vp, P(u32), P(f32), P(f32)]),
        "wpt_ctx_photon_tree": (C.c_int64, [vp, P(u32), P(f32), P(f32)]),
        "wpt_ctx_photon_sample": (i32, [vp, P(f32), P(u32), u64, P(u32), P(f32)]),
        "wpt_ctx_error_map": (i32, [vp, P(f32), P(f32)]),
        "wpt_ctx_round_spp": (i32, [vp, P(u32)]),
        "wpt_ctx_device_buffers": (i32, [vp, P(u64), P(u64)]),
        "wpt_ctx_set_stream": (i32, [vp, u64]),
        "wpt_ctx_upload_scene": (C.c_int64, [vp]),
        "wpt_ctx_profile": (i32, [vp, i32]),
        "wpt_ctx_profile_read": (i32, [vp, P(C.c_double)]),
        "wpt_ctx_profile_read_rounds": (i32, [vp, P(C.c_double)]),
        "wpt_ctx_mark_accum_dirty": (i32, [vp]),
        "wpt_ctx_load_obj": (C.c_int64, [vp, u32, C.c_char_p, i32]),
        "wpt_parse_obj": (C.c_int64, [C.c_char_p, u64, i32, P(f32), u64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._wpt_symbols = sorted(sig)
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def parse_obj(text, client_scale=True):
    """OBJ text -> (n, 3) float32 vertices, 3 per triangle (obj_parser.ts:3-51, index.ts:216-220)."""
    L = load_library()
    b = text.encode() if isinstance(text, str) else bytes(text)
    n = L.wpt_parse_obj(b, len(b), int(client_scale), None, 0)
    if n < 0:
        raise WptError(L.wpt_last_error().decode())
    out = np.empty(n, np.float32)
    L.wpt_parse_obj(b, len(b), int(client_scale), _p(out, C.c_float), n)
    return out.reshape(-1, 3)


def nccl_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the host sends it to the other ranks)."""
    L = load_library()
    buf = (C.c_uint8 * 128)()
    if L.wpt_nccl_unique_id(buf) != 0:
        raise WptError(L.wpt_last_error().decode())
    return bytes(buf)


class PathTracer:
    """One rendering session = the reference's `Config` (wasm_interface.rs:37-57) on one GPU."""

    def __init__(self, width, height, scene_id, cam_x, cam_y, cam_z, cam_rot_x, cam_rot_y, device=-1):
        """`init(width, height, scene_id, cam_x, cam_y, cam_z, cam_rot_x, cam_rot_y)`, wasm_interface.rs:67-113."""
        self.L = load_library()
        self.W, self.H = int(width), int(height)
        self.h = self.L.wpt_ctx_create(device, width, height, scene_id, cam_x, cam_y, cam_z, cam_rot_x, cam_rot_y)
        if not self.h:
            raise WptError(self.L.wpt_last_error().decode())
        self.h = C.c_void_p(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.wpt_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc is None or rc < 0:
            raise WptError(self.L.wpt_last_error().decode())
        return rc

    # ------------------------------------------------------------ reference surface
    def results(self, is_show_sampling=0):
        """wasm_interface.rs:120-134 — (H, W, 4) uint8 view of the library-owned host buffer."""
        ptr = self.L.wpt_ctx_results(self.h, is_show_sampling)
        if not ptr:
            raise WptError(self.L.wpt_last_error().decode())
        return np.ctypeslib.as_array(ptr, shape=(self.H, self.W, 4))

    def reset(self):
        self._chk(self.L.wpt_ctx_reset(self.h))

    def update_scene(self, scene_id):
        self._chk(self.L.wpt_ctx_update_scene(self.h, scene_id))

    def update_settings(self, left_type, right_type, is_left_adaptive, is_right_adaptive, is_light_debug):
        self._chk(self.L.wpt_ctx_update_settings(self.h, left_type, right_type, is_left_adaptive, is_right_adaptive, is_light_debug))

    def update_viewport(self, width, height):
        self._chk(self.L.wpt_ctx_update_viewport(self.h, width, height))
        self.W, self.H = int(width), int(height)

    def update_camera(self, cam_x, cam_y, cam_z, cam_rot_x, cam_rot_y):
        self._chk(self.L.wpt_ctx_update_camera(self.h, cam_x, cam_y, cam_z, cam_rot_x, cam_rot_y))

    def allocate_mesh(self, mesh_id, num_vertices):
        self._chk(self.L.wpt_ctx_allocate_mesh(self.h, mesh_id, num_vertices))

    def mesh_vertices(self, mesh_id, num_vertices):
        """wasm_interface.rs:275-288 — writable (num_vertices, 3) float32 view of library memory."""
        ptr = self.L.wpt_ctx_mesh_vertices(self.h, mesh_id)
        if not ptr:
            raise WptError(self.L.wpt_last_error().decode())
        return np.ctypeslib.as_array(ptr, shape=(num_vertices, 3))

    def notify_mesh_loaded(self, mesh_id):
        return self._chk(self.L.wpt_ctx_notify_mesh_loaded(self.h, mesh_id)) == 1

    def allocate_texture(self, tex_id, width, height):
        ptr = self.L.wpt_ctx_allocate_texture(self.h, tex_id, width, height)
        if not ptr:
            raise WptError(self.L.wpt_last_error().decode())
        return np.ctypeslib.as_array(ptr, shape=(height, width, 3))

    def notify_texture_loaded(self, tex_id):
        return self._chk(self.L.wpt_ctx_notify_texture_loaded(self.h, tex_id)) == 1

    def store_texture(self, tex_id, rgb):
        """The worker's texture upload (src_ts/worker/worker.ts:182-190): rgb = uint8 array (height, width, 3)."""
        t = np.ascontiguousarray(rgb, np.uint8)
        self.allocate_texture(tex_id, t.shape[1], t.shape[0])[:] = t
        return self.notify_texture_loaded(tex_id)

    def compute(self, num_samples):
        self._chk(self.L.wpt_ctx_compute(self.h, num_samples))

    def store_mesh(self, mesh_id, vertices):
        """The worker's three-step upload (src_ts/worker/worker.ts:171-179)."""
        v = np.ascontiguousarray(vertices, np.float32).reshape(-1, 3)
        self.allocate_mesh(mesh_id, len(v))
        self.mesh_vertices(mesh_id, len(v))[:] = v
        return self.notify_mesh_loaded(mesh_id)

    # ------------------------------------------------------------ additions (wpt.h section 3)
    def get_config(self):
        cfg = WptConfig()
        self._chk(self.L.wpt_ctx_get_config(self.h, C.byref(cfg)))
        return cfg

    def set_config(self, **kw):
        cfg = self.get_config()
        for k, v in kw.items():
            if not hasattr(cfg, k):
                raise WptError("unknown config field " + k)
            setattr(cfg, k, v)
        self._chk(self.L.wpt_ctx_set_config(self.h, C.byref(cfg)))

    def render_exact(self, spp):
        self._chk(self.L.wpt_ctx_render_exact(self.h, spp))

    def render_adaptive(self, budget_ticks):
        return self._chk(self.L.wpt_ctx_render_adaptive(self.h, budget_ticks))

    def render_random(self, ticks):
        self._chk(self.L.wpt_ctx_render_random(self.h, ticks))

    def set_exchange_callback(self, fn):
        """fn() is called between adaptive rounds (multi-GPU accumulator exchange); None removes it."""
        if fn is None:
            self._cb = None
            self._chk(self.L.wpt_ctx_set_exchange_callback(self.h, None, None))
            return
        self._cb = C.CFUNCTYPE(None, C.c_void_p)(lambda _u: fn())
        self._chk(self.L.wpt_ctx_set_exchange_callback(self.h, C.cast(self._cb, C.c_void_p), None))

    def set_reduce_callback(self, fn):
        """fn(dev_ptr, n_words) must sum `n_words` uint32 at `dev_ptr` over all ranks in place, on the session's
        stream (multi-GPU photon warm-up: each rank emits every world-th shot); None removes it."""
        if fn is None:
            self._rcb = None
            self._chk(self.L.wpt_ctx_set_reduce_callback(self.h, None, None))
            return
        self._rcb = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_uint64)(lambda _u, p, n: fn(int(p), int(n)))
        self._chk(self.L.wpt_ctx_set_reduce_callback(self.h, C.cast(self._rcb, C.c_void_p), None))

    def attach_nccl(self, unique_id, rank, world):
        """Native multi-GPU plane: `unique_id` = the 128 bytes of nccl_unique_id() of rank 0 (None when world == 1)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id)) if unique_id is not None else None
        self._chk(self.L.wpt_ctx_attach_nccl(self.h, buf, rank, world))

    def detach_nccl(self):
        self._chk(self.L.wpt_ctx_detach_nccl(self.h))

    def gather_frame(self):
        self._chk(self.L.wpt_ctx_gather_frame(self.h))

    def build_photons(self):
        self._chk(self.L.wpt_ctx_build_photons(self.h))

    def synchronize(self):
        self._chk(self.L.wpt_ctx_synchronize(self.h))

    def stats(self):
        out = np.zeros(8, np.uint64)
        self._chk(self.L.wpt_ctx_stats(self.h, _p(out, C.c_uint64)))
        k = ["rays", "paths", "node_visits", "photons_shot", "photons_stored", "iterations", "launches"]
        return {a: int(b) for a, b in zip(k, out)}

    def primary_probe(self):
        n = self.W * self.H
        ids = np.empty(n, np.int32); vis = np.empty(n, np.uint32); dist = np.empty(n, np.float32)
        self._chk(self.L.wpt_ctx_primary_probe(self.h, _p(ids, C.c_int32), _p(vis, C.c_uint32), _p(dist, C.c_float)))
        return ids.reshape(self.H, self.W), vis.reshape(self.H, self.W), dist.reshape(self.H, self.W)

    def accum(self):
        n = self.W * self.H
        rgb = np.empty(n * 3, np.float32); cnt = np.empty(n, np.uint32)
        self._chk(self.L.wpt_ctx_accum(self.h, _p(rgb, C.c_float), _p(cnt, C.c_uint32)))
        return rgb.reshape(self.H, self.W, 3), cnt.reshape(self.H, self.W)

    def trace_rays(self, origins, dirs, want_normals=True):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3); d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = len(o)
        ids = np.empty(n, np.int32); dist = np.empty(n, np.float32); vis = np.empty(n, np.uint32)
        nrm = np.empty((n, 3), np.float32) if want_normals else None
        self._chk(self.L.wpt_ctx_trace_rays(self.h, _p(o, C.c_float), _p(d, C.c_float), n, _p(ids, C.c_int32), _p(dist, C.c_float), _p(vis, C.c_uint32), _p(nrm, C.c_float)))
        return ids, dist, vis, nrm

    def scene_info(self):
        out = np.zeros(8, np.uint64)
        self._chk(self.L.wpt_ctx_scene_info(self.h, _p(out, C.c_uint64)))
        k = ["num_shapes", "num_inf", "num_lights", "bvh2_nodes", "bvh2_depth", "bvh4_nodes", "bvh_kind", "bvh4_depth"]
        return {a: int(b) for a, b in zip(k, out)}

    def bvh2(self):
        n = self.scene_info()["bvh2_nodes"]
        b = np.empty(n * 6, np.float32); lf = np.empty(n, np.uint32); cnt = np.empty(n, np.uint32)
        self._chk(self.L.wpt_ctx_bvh2(self.h, _p(b, C.c_float), _p(lf, C.c_uint32), _p(cnt, C.c_uint32)))
        return b.reshape(n, 6), lf, cnt

    def bvh4(self):
        n = self.scene_info()["bvh4_nodes"]
        b = np.empty(n * 24, np.float32); ch = np.empty(n * 4, np.int32); nc = np.empty(n, np.uint32)
        self._chk(self.L.wpt_ctx_bvh4(self.h, _p(b, C.c_float), _p(ch, C.c_int32), _p(nc, C.c_uint32)))
        return b.reshape(n, 4, 6), ch.reshape(n, 4), nc

    def shape_order(self):
        n = self.scene_info()["num_shapes"]
        src = np.empty(n, np.int32); typ = np.empty(n, np.int32)
        self._chk(self.L.wpt_ctx_shape_order(self.h, _p(src, C.c_int32), _p(typ, C.c_int32)))
        return src, typ

    def lights(self):
        n = self.scene_info()["num_lights"]
        out = np.empty(n, np.uint32)
        self._chk(self.L.wpt_ctx_lights(self.h, _p(out, C.c_uint32)))
        return out

    def photons(self):
        shots = C.c_uint64(0)
        n = self._chk(self.L.wpt_ctx_photon_count(self.h, C.byref(shots)))
        light = np.empty(n, np.uint32); loc = np.empty(n * 3, np.float32); w = np.empty(n, np.float32)
        self._chk(self.L.wpt_ctx_photon_list(self.h, _p(light, C.c_uint32), _p(loc, C.c_float), _p(w, C.c_float)))
        return light, loc.reshape(n, 3), w, int(shots.value)

    def photon_tree(self):
        n = self._chk(self.L.wpt_ctx_photon_tree(self.h, None, None, None))
        nl = self.scene_info()["num_lights"]
        meta = np.empty(n * 3, np.uint32); cum = np.empty(n * nl, np.float32); bins = np.empty(n * nl, np.float32)
        self._chk(self.L.wpt_ctx_photon_tree(self.h, _p(meta, C.c_uint32), _p(cum, C.c_float), _p(bins, C.c_float)))
        return meta.reshape(n, 3), cum.reshape(n, nl), bins.reshape(n, nl)

    def photon_sample(self, pts, seeds):
        p = np.ascontiguousarray(pts, np.float32).reshape(-1, 3); s = np.ascontiguousarray(seeds, np.uint32)
        n = len(p)
        light = np.empty(n, np.uint32); pdf = np.empty(n, np.float32)
        self._chk(self.L.wpt_ctx_photon_sample(self.h, _p(p, C.c_float), _p(s, C.c_uint32), n, _p(light, C.c_uint32), _p(pdf, C.c_float)))
        return light, pdf

    def error_map(self):
        cfg = self.get_config()
        rw, rh = cfg.region_w or self.W, cfg.region_h or self.H
        mse = np.empty(rw * rh, np.float32); st = np.empty(3, np.float32)
        self._chk(self.L.wpt_ctx_error_map(self.h, _p(mse, C.c_float), _p(st, C.c_float)))
        return mse.reshape(rh, rw), st

    def round_spp(self):
        cfg = self.get_config()
        rw, rh = cfg.region_w or self.W, cfg.region_h or self.H
        out = np.empty(rw * rh, np.uint32)
        self._chk(self.L.wpt_ctx_round_spp(self.h, _p(out, C.c_uint32)))
        return out.reshape(rh, rw)

    def device_buffers(self):
        ptrs = np.zeros(8, np.uint64); sizes = np.zeros(8, np.uint64)
        self._chk(self.L.wpt_ctx_device_buffers(self.h, _p(ptrs, C.c_uint64), _p(sizes, C.c_uint64)))
        return dict(accum=(int(ptrs[0]), int(sizes[0])), rgba=(int(ptrs[1]), int(sizes[1])), sampling=(int(ptrs[2]), int(sizes[2])), stream=int(ptrs[3]))

    def set_stream(self, cuda_stream):
        self._chk(self.L.wpt_ctx_set_stream(self.h, int(cuda_stream)))

    def upload_scene(self):
        return self._chk(self.L.wpt_ctx_upload_scene(self.h))

    def profile(self, enable=True):
        self._chk(self.L.wpt_ctx_profile(self.h, int(enable)))

    def profile_read(self):
        out = np.zeros(8, np.float64)
        self._chk(self.L.wpt_ctx_profile_read(self.h, _p(out, C.c_double)))
        return dict(trace_ms=out[0], trace_launches=int(out[1]), shade_ms=out[2], shade_launches=int(out[3]),
                    prim_tests=int(out[4]), rays=int(out[5]), node_visits=int(out[6]))

    def profile_read_rounds(self):
        out = np.zeros(8, np.float64)
        self._chk(self.L.wpt_ctx_profile_read_rounds(self.h, _p(out, C.c_double)))
        return dict(rounds=int(out[0]), error_map_ms=out[1], render_ms=out[2], exchange_ms=out[3], photon_warmup_ms=out[4], collectives=int(out[5]))

    def mark_accum_dirty(self):
        self._chk(self.L.wpt_ctx_mark_accum_dirty(self.h))

    def load_obj(self, mesh_id, path, apply_client_scale=True):
        return self._chk(self.L.wpt_ctx_load_obj(self.h, mesh_id, os.fsencode(path), int(apply_client_scale)))
