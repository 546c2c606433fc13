//! Safe wrapper over the handle API of `libwpt.so`. Method names follow the reference's exports
//! (`src/wasm_interface.rs:65-384`). NOT COMPILED OR TESTED HERE (no Rust toolchain).
use std::ffi::{CStr, CString};
use wpt_sys as sys;

#[derive(Debug)]
pub struct Error(pub String);
pub type Result<T> = std::result::Result<T, Error>;

fn last_error() -> Error {
    unsafe { Error(CStr::from_ptr(sys::wpt_last_error()).to_string_lossy().into_owned()) }
}
fn check(rc: i32) -> Result<()> { if rc < 0 { Err(last_error()) } else { Ok(()) } }

#[derive(Clone, Copy, Debug, PartialEq)]
pub enum RenderType { NoNEE = 0, NormalNEE = 1, PNEE = 2 }

/// One rendering session on one GPU = the reference's `Config` (wasm_interface.rs:37-57).
pub struct PathTracer { ctx: *mut sys::wpt_ctx, width: u32, height: u32 }

impl PathTracer {
    /// `init(width, height, scene_id, cam_x, cam_y, cam_z, cam_rot_x, cam_rot_y)`
    pub fn new(device: i32, width: u32, height: u32, scene_id: u32, cam: [f32; 5]) -> Result<Self> {
        let ctx = unsafe { sys::wpt_ctx_create(device, width, height, scene_id, cam[0], cam[1], cam[2], cam[3], cam[4]) };
        if ctx.is_null() { Err(last_error()) } else { Ok(PathTracer { ctx, width, height }) }
    }
    /// `results(is_show_sampling)`: RGBA8, row-major, borrowed until the next `update_viewport`.
    pub fn results(&mut self, is_show_sampling: bool) -> Result<&[u8]> {
        let p = unsafe { sys::wpt_ctx_results(self.ctx, is_show_sampling as u32) };
        if p.is_null() { return Err(last_error()); }
        Ok(unsafe { std::slice::from_raw_parts(p, (self.width * self.height * 4) as usize) })
    }
    pub fn reset(&mut self) -> Result<()> { check(unsafe { sys::wpt_ctx_reset(self.ctx) }) }
    pub fn update_scene(&mut self, scene_id: u32) -> Result<()> { check(unsafe { sys::wpt_ctx_update_scene(self.ctx, scene_id) }) }
    pub fn update_settings(&mut self, left: RenderType, right: RenderType, left_adaptive: bool, right_adaptive: bool, light_debug: bool) -> Result<()> {
        check(unsafe { sys::wpt_ctx_update_settings(self.ctx, left as u32, right as u32, left_adaptive as u32, right_adaptive as u32, light_debug as u32) })
    }
    pub fn update_viewport(&mut self, width: u32, height: u32) -> Result<()> {
        check(unsafe { sys::wpt_ctx_update_viewport(self.ctx, width, height) })?;
        self.width = width; self.height = height;
        Ok(())
    }
    pub fn update_camera(&mut self, cam: [f32; 5]) -> Result<()> {
        check(unsafe { sys::wpt_ctx_update_camera(self.ctx, cam[0], cam[1], cam[2], cam[3], cam[4]) })
    }
    /// allocate_mesh + mesh_vertices + notify_mesh_loaded (worker.ts:171-179); `vertices` = 9 floats per triangle.
    pub fn store_mesh(&mut self, id: u32, vertices: &[f32]) -> Result<bool> {
        let nv = (vertices.len() / 3) as u32;
        check(unsafe { sys::wpt_ctx_allocate_mesh(self.ctx, id, nv) })?;
        let dst = unsafe { sys::wpt_ctx_mesh_vertices(self.ctx, id) };
        if dst.is_null() { return Err(last_error()); }
        unsafe { std::ptr::copy_nonoverlapping(vertices.as_ptr(), dst, (nv * 3) as usize) };
        let rc = unsafe { sys::wpt_ctx_notify_mesh_loaded(self.ctx, id) };
        check(rc)?;
        Ok(rc == 1)
    }
    pub fn load_obj(&mut self, id: u32, path: &str, apply_client_scale: bool) -> Result<i64> {
        let c = CString::new(path).map_err(|e| Error(e.to_string()))?;
        let n = unsafe { sys::wpt_ctx_load_obj(self.ctx, id, c.as_ptr(), apply_client_scale as i32) };
        if n < 0 { Err(last_error()) } else { Ok(n) }
    }
    /// `compute(num_samples)`
    pub fn compute(&mut self, num_samples: u64) -> Result<()> { check(unsafe { sys::wpt_ctx_compute(self.ctx, num_samples) }) }
    pub fn config(&self) -> Result<sys::wpt_config> {
        let mut c = sys::wpt_config::default();
        check(unsafe { sys::wpt_ctx_get_config(self.ctx, &mut c) })?;
        Ok(c)
    }
    pub fn set_config(&mut self, cfg: &sys::wpt_config) -> Result<()> { check(unsafe { sys::wpt_ctx_set_config(self.ctx, cfg) }) }
    pub fn render_exact(&mut self, spp: u32) -> Result<()> { check(unsafe { sys::wpt_ctx_render_exact(self.ctx, spp) }) }
    pub fn render_adaptive(&mut self, budget_ticks: u64) -> Result<u64> {
        let n = unsafe { sys::wpt_ctx_render_adaptive(self.ctx, budget_ticks) };
        if n < 0 { Err(last_error()) } else { Ok(n as u64) }
    }
    pub fn build_photons(&mut self) -> Result<()> { check(unsafe { sys::wpt_ctx_build_photons(self.ctx) }) }
    /// Native multi-GPU plane: join the job described by rank 0's `nccl_unique_id()` (the host distributes the 128 bytes).
    /// Sets rank / world (4-row bands, band b belongs to rank b % world); the library then all-gathers the accumulator rows
    /// between adaptive rounds and on `gather_frame`, and merges the photon batches with an integer allreduce.
    pub fn attach_nccl(&mut self, id: &[u8; 128], rank: u32, world: u32) -> Result<()> {
        check(unsafe { sys::wpt_ctx_attach_nccl(self.ctx, id.as_ptr(), rank, world) })
    }
    pub fn detach_nccl(&mut self) -> Result<()> { check(unsafe { sys::wpt_ctx_detach_nccl(self.ctx) }) }
    pub fn gather_frame(&mut self) -> Result<()> { check(unsafe { sys::wpt_ctx_gather_frame(self.ctx) }) }
    /// rays, paths, BVH node visits (the reference's `num_bvh_hits`), photons shot, photons stored, iterations, launches
    pub fn stats(&mut self) -> Result<[u64; 8]> {
        let mut out = [0u64; 8];
        check(unsafe { sys::wpt_ctx_stats(self.ctx, out.as_mut_ptr()) })?;
        Ok(out)
    }
}

/// `ncclGetUniqueId` through the library (rank 0 calls it; every rank passes the bytes to `PathTracer::attach_nccl`).
pub fn nccl_unique_id() -> Result<[u8; 128]> {
    let mut id = [0u8; 128];
    check(unsafe { sys::wpt_nccl_unique_id(id.as_mut_ptr()) })?;
    Ok(id)
}

impl Drop for PathTracer {
    fn drop(&mut self) { unsafe { sys::wpt_ctx_destroy(self.ctx) } }
}
