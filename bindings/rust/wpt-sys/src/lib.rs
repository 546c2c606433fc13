//! Raw bindings to `include/wpt.h`. NOT COMPILED OR TESTED HERE (no Rust toolchain in the build
//! environment); kept in sync with the header by hand. Each function replaces the export of
//! `src/wasm_interface.rs` named in `wpt.h`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct wpt_ctx { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct wpt_config {
    pub bvh_kind: u32,
    pub render_type: u32,
    pub light_debug: u32,
    pub base_seed: u32,
    pub photon_target: u64,
    pub region_x: u32,
    pub region_y: u32,
    pub region_w: u32,
    pub region_h: u32,
    pub rank: u32,
    pub world: u32,
    pub engine: u32,
    pub reserved: [u32; 3],
}

pub const WPT_NO_NEE: u32 = 0;
pub const WPT_NORMAL_NEE: u32 = 1;
pub const WPT_PNEE: u32 = 2;
pub const WPT_SCENE_MUSEUM: u32 = 0;
pub const WPT_SCENE_BUNNY: u32 = 2;
/// Extension scene: not a reference id.
pub const WPT_SCENE_EXT_WHITTED: u32 = 256;
pub const WPT_DEVICE_NONE: c_int = -2;

extern "C" {
    // global instance = the reference's `static mut CONFIG`
    pub fn wpt_init(width: u32, height: u32, scene_id: u32, cam_x: f32, cam_y: f32, cam_z: f32, cam_rot_x: f32, cam_rot_y: f32);
    pub fn wpt_results(is_show_sampling: u32) -> *const u8;
    pub fn wpt_reset();
    pub fn wpt_update_scene(scene_id: u32);
    pub fn wpt_update_settings(left_type: u32, right_type: u32, is_left_adaptive: u32, is_right_adaptive: u32, is_light_debug: u32);
    pub fn wpt_update_viewport(width: u32, height: u32);
    pub fn wpt_update_camera(cam_x: f32, cam_y: f32, cam_z: f32, cam_rot_x: f32, cam_rot_y: f32);
    pub fn wpt_allocate_mesh(id: u32, num_vertices: u32);
    pub fn wpt_mesh_vertices(id: u32) -> *mut f32;
    pub fn wpt_notify_mesh_loaded(id: u32) -> c_int;
    pub fn wpt_allocate_texture(id: u32, width: u32, height: u32) -> *mut u8;
    pub fn wpt_notify_texture_loaded(id: u32) -> c_int;
    pub fn wpt_compute(num_samples: u64);
    pub fn wpt_last_error() -> *const c_char;
    pub fn wpt_global_ctx() -> *mut wpt_ctx;

    // handle API
    pub fn wpt_ctx_create(device: c_int, width: u32, height: u32, scene_id: u32, cam_x: f32, cam_y: f32, cam_z: f32, cam_rot_x: f32, cam_rot_y: f32) -> *mut wpt_ctx;
    pub fn wpt_ctx_destroy(ctx: *mut wpt_ctx);
    pub fn wpt_ctx_results(ctx: *mut wpt_ctx, is_show_sampling: u32) -> *const u8;
    pub fn wpt_ctx_reset(ctx: *mut wpt_ctx) -> c_int;
    pub fn wpt_ctx_update_scene(ctx: *mut wpt_ctx, scene_id: u32) -> c_int;
    pub fn wpt_ctx_update_settings(ctx: *mut wpt_ctx, left_type: u32, right_type: u32, is_left_adaptive: u32, is_right_adaptive: u32, is_light_debug: u32) -> c_int;
    pub fn wpt_ctx_update_viewport(ctx: *mut wpt_ctx, width: u32, height: u32) -> c_int;
    pub fn wpt_ctx_update_camera(ctx: *mut wpt_ctx, cam_x: f32, cam_y: f32, cam_z: f32, cam_rot_x: f32, cam_rot_y: f32) -> c_int;
    pub fn wpt_ctx_allocate_mesh(ctx: *mut wpt_ctx, id: u32, num_vertices: u32) -> c_int;
    pub fn wpt_ctx_mesh_vertices(ctx: *mut wpt_ctx, id: u32) -> *mut f32;
    pub fn wpt_ctx_notify_mesh_loaded(ctx: *mut wpt_ctx, id: u32) -> c_int;
    pub fn wpt_ctx_compute(ctx: *mut wpt_ctx, num_samples: u64) -> c_int;

    // additions
    pub fn wpt_default_config(cfg: *mut wpt_config);
    pub fn wpt_ctx_set_config(ctx: *mut wpt_ctx, cfg: *const wpt_config) -> c_int;
    pub fn wpt_ctx_get_config(ctx: *mut wpt_ctx, cfg: *mut wpt_config) -> c_int;
    pub fn wpt_ctx_render_exact(ctx: *mut wpt_ctx, spp: u32) -> c_int;
    pub fn wpt_ctx_render_adaptive(ctx: *mut wpt_ctx, budget_ticks: u64) -> i64;
    pub fn wpt_ctx_render_random(ctx: *mut wpt_ctx, ticks: u64) -> c_int;
    pub fn wpt_ctx_allocate_texture(ctx: *mut wpt_ctx, id: u32, width: u32, height: u32) -> *mut u8;
    pub fn wpt_ctx_notify_texture_loaded(ctx: *mut wpt_ctx, id: u32) -> c_int;
    /// Multi-GPU: called between adaptive rounds to exchange the accumulator rows (NCCL all-gather in the caller).
    pub fn wpt_ctx_set_exchange_callback(ctx: *mut wpt_ctx, callback: Option<unsafe extern "C" fn(user: *mut c_void)>, user: *mut c_void) -> c_int;
    /// Multi-GPU photon warm-up: in-place integer sum of `n_words` u32 at `dev_words` over all ranks (ncclAllReduce, ncclUint32, ncclSum).
    // native multi-GPU plane (NCCL inside libwpt, csrc/dist_nccl.cpp)
    pub fn wpt_nccl_unique_id(out: *mut u8) -> c_int;
    pub fn wpt_ctx_attach_nccl(ctx: *mut wpt_ctx, id: *const u8, rank: u32, world: u32) -> c_int;
    pub fn wpt_ctx_attach_nccl_comm(ctx: *mut wpt_ctx, nccl_comm: *mut c_void, rank: u32, world: u32) -> c_int;
    pub fn wpt_ctx_detach_nccl(ctx: *mut wpt_ctx) -> c_int;
    pub fn wpt_ctx_gather_frame(ctx: *mut wpt_ctx) -> c_int;
    pub fn wpt_ctx_profile_read_rounds(ctx: *mut wpt_ctx, out: *mut f64) -> c_int;
    pub fn wpt_ctx_set_reduce_callback(ctx: *mut wpt_ctx, callback: Option<unsafe extern "C" fn(user: *mut c_void, dev_words: *mut c_void, n_words: u64)>, user: *mut c_void) -> c_int;
    pub fn wpt_ctx_device_buffers(ctx: *mut wpt_ctx, ptrs: *mut u64, sizes: *mut u64) -> c_int;
    pub fn wpt_ctx_mark_accum_dirty(ctx: *mut wpt_ctx) -> c_int;
    pub fn wpt_ctx_build_photons(ctx: *mut wpt_ctx) -> c_int;
    pub fn wpt_ctx_synchronize(ctx: *mut wpt_ctx) -> c_int;
    pub fn wpt_ctx_stats(ctx: *mut wpt_ctx, out: *mut u64) -> c_int;
    pub fn wpt_ctx_primary_probe(ctx: *mut wpt_ctx, ids: *mut i32, visits: *mut u32, dist: *mut f32) -> c_int;
    pub fn wpt_ctx_accum(ctx: *mut wpt_ctx, rgb: *mut f32, counts: *mut u32) -> c_int;
    pub fn wpt_ctx_load_obj(ctx: *mut wpt_ctx, id: u32, path: *const c_char, apply_client_scale: c_int) -> i64;
}
