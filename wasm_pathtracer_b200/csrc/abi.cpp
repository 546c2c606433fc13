// extern "C" surface of libwpt.so — see include/wpt.h for the reference citations.
#include "context.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

using namespace wpt;

static thread_local std::string g_err;
static Context* g_ctx = nullptr;   // the reference's `static mut CONFIG` (wasm_interface.rs:62)

static bool strict_mode() { const char* s = std::getenv("WPT_STRICT"); return s && s[0] == '1'; }
static void fail(const char* what) {
  g_err = what;
  if (strict_mode()) { std::fprintf(stderr, "wpt: %s\n", what); std::abort(); }   // the reference traps
}
// Every entry point runs on the session's own device: C() selects it (allocations, pinned memory and launches of a
// session must not land on whatever device the calling thread happened to use last) and the guard puts the caller's
// device back afterwards.
struct DeviceRestore {
  int prev = -1; bool have = false;
  DeviceRestore() { have = cudaGetDevice(&prev) == cudaSuccess; if (!have) cudaGetLastError(); }
  ~DeviceRestore() { int now = -1; if (have && cudaGetDevice(&now) == cudaSuccess && now != prev) cudaSetDevice(prev); }
};
template <class F> static int guard(F&& f) {
  DeviceRestore restore;
  try { f(); return 0; }
  catch (const std::exception& e) { fail(e.what()); return -1; }
  catch (...) { fail("unknown error"); return -1; }
}
static Context* C(wpt_ctx* c) {
  if (!c) throw std::runtime_error("init not called");
  Context* x = reinterpret_cast<Context*>(c);
  if (x->has_device) WPT_CUDA(cudaSetDevice(x->device));
  return x;
}

extern "C" {

const char* wpt_last_error(void) { return g_err.c_str(); }
wpt_ctx* wpt_global_ctx(void) { return reinterpret_cast<wpt_ctx*>(g_ctx); }

void wpt_default_config(wpt_config* cfg) {
  std::memset(cfg, 0, sizeof *cfg);
  cfg->bvh_kind = 2;                 // scene.rs:60
  cfg->render_type = WPT_NORMAL_NEE;
  cfg->base_seed = 0xBABABEBEu;      // rng.rs:11
  cfg->photon_target = 300000;       // tracer.rs:104
  cfg->world = 1;
}

// ------------------------------------------------------------------ handle API
wpt_ctx* wpt_ctx_create(int device, uint32_t width, uint32_t height, uint32_t scene_id, float cx, float cy, float cz, float rx, float ry) {
  Context* c = nullptr;
  float cam[5] = {cx, cy, cz, rx, ry};
  if (guard([&] {
        if (width == 0 || height == 0) throw std::runtime_error("empty viewport");
        c = new Context(device, width, height, scene_id, cam);
      }) != 0) return nullptr;
  return reinterpret_cast<wpt_ctx*>(c);
}
void wpt_ctx_destroy(wpt_ctx* ctx) { DeviceRestore restore; if (ctx == reinterpret_cast<wpt_ctx*>(g_ctx)) g_ctx = nullptr; delete reinterpret_cast<Context*>(ctx); }

const uint8_t* wpt_ctx_results(wpt_ctx* ctx, uint32_t show) {
  const uint8_t* p = nullptr;
  guard([&] { p = C(ctx)->results(show); });
  return p;
}
int wpt_ctx_reset(wpt_ctx* ctx) { return guard([&] { C(ctx)->reset(); }); }
int wpt_ctx_update_scene(wpt_ctx* ctx, uint32_t id) {   // wasm_interface.rs:154-168
  return guard([&] { Context* c = C(ctx); c->select_scene(id); c->reset(); });
}
int wpt_ctx_update_settings(wpt_ctx* ctx, uint32_t lt, uint32_t rt, uint32_t la, uint32_t ra, uint32_t dbg) {   // wasm_interface.rs:173-214
  return guard([&] {
    Context* c = C(ctx);
    if (lt > 2 || rt > 2) throw std::runtime_error("Invalid RenderType magic number");
    c->left = HalfSettings{lt, la == 1 ? 1u : 0u}; c->right = HalfSettings{rt, ra == 1 ? 1u : 0u};
    c->light_debug = dbg == 1; c->cfg.light_debug = c->light_debug;
    c->photons_ready = false;   // fresh RenderInstances discard their photons (wasm_interface.rs:198-199)
    c->reset();
  });
}
int wpt_ctx_update_viewport(wpt_ctx* ctx, uint32_t w, uint32_t h) {   // wasm_interface.rs:219-232
  return guard([&] {
    Context* c = C(ctx);
    if (w == 0 || h == 0) throw std::runtime_error("empty viewport");
    if (c->has_device) WPT_CUDA(cudaStreamSynchronize(c->stream));
    c->W = w; c->H = h;
    c->cfg.region_x = c->cfg.region_y = c->cfg.region_w = c->cfg.region_h = 0;
    c->alloc_targets();
    c->reset();
  });
}
int wpt_ctx_update_camera(wpt_ctx* ctx, float x, float y, float z, float rx, float ry) {   // wasm_interface.rs:239-248
  return guard([&] { Context* c = C(ctx); float cam[5] = {x, y, z, rx, ry}; std::memcpy(c->cam, cam, sizeof cam); c->reset(); });
}
int wpt_ctx_allocate_mesh(wpt_ctx* ctx, uint32_t id, uint32_t nv) {   // wasm_interface.rs:259-270
  return guard([&] { Context* c = C(ctx); c->mesh_tris.erase(id); c->mesh_preload[id].assign((size_t)nv * 3, 0.0f); });
}
float* wpt_ctx_mesh_vertices(wpt_ctx* ctx, uint32_t id) {   // wasm_interface.rs:275-288
  float* p = nullptr;
  guard([&] {
    Context* c = C(ctx);
    auto it = c->mesh_preload.find(id);
    if (it == c->mesh_preload.end()) throw std::runtime_error("Mesh not allocated");
    p = it->second.data();
  });
  return p;
}
int wpt_ctx_notify_mesh_loaded(wpt_ctx* ctx, uint32_t id) {   // wasm_interface.rs:293-329
  int res = 0;
  if (guard([&] {
        Context* c = C(ctx);
        auto it = c->mesh_preload.find(id);
        if (it != c->mesh_preload.end()) {
          c->mesh_tris[id] = mesh_triangles(it->second.data(), it->second.size() / 3, 2);
          c->mesh_preload.erase(it);
        }
        if ((id == 0 && c->scene_id == 1) || (id == 1 && c->scene_id == 2) || (id == 2 && c->scene_id == 3)) {
          c->select_scene(c->scene_id);
          c->reset();
          res = 1;
        }
      }) != 0) return -1;
  return res;
}
uint8_t* wpt_ctx_allocate_texture(wpt_ctx* ctx, uint32_t id, uint32_t w, uint32_t h) {   // wasm_interface.rs:335-352
  uint8_t* p = nullptr;
  guard([&] { Context* c = C(ctx); c->textures[id].assign((size_t)w * h * 3 + 1, 0); c->tex_dims[id] = std::make_pair(w, h); p = c->textures[id].data(); });
  return p;
}
int wpt_ctx_notify_texture_loaded(wpt_ctx* ctx, uint32_t) {   // wasm_interface.rs:357-366: stub, always false
  if (guard([&] { C(ctx); }) != 0) return -1;
  return 0;
}
int wpt_ctx_compute(wpt_ctx* ctx, uint64_t n) { return guard([&] { C(ctx)->compute(n); }); }

int wpt_ctx_set_config(wpt_ctx* ctx, const wpt_config* cfg) {
  return guard([&] {
    Context* c = C(ctx);
    if (cfg->bvh_kind != 2 && cfg->bvh_kind != 4) throw std::runtime_error("bvh_kind must be 2 or 4");
    if (cfg->render_type > 2) throw std::runtime_error("Invalid RenderType magic number");
    if (cfg->world == 0 || cfg->rank >= cfg->world) throw std::runtime_error("invalid rank/world");
    bool rebuild = cfg->bvh_kind != c->cfg.bvh_kind;
    bool rephoton = cfg->base_seed != c->cfg.base_seed || cfg->photon_target != c->cfg.photon_target;
    wpt_config old = c->cfg;
    c->cfg = *cfg;
    if (rebuild) {
      try { c->select_scene(c->scene_id); }
      catch (...) { c->cfg = old; throw; }   // e.g. BVH4 of a scene whose root is a leaf (bvh4.rs:67)
    }
    c->light_debug = cfg->light_debug;
    if (rephoton) c->photons_ready = false;
  });
}
int wpt_ctx_get_config(wpt_ctx* ctx, wpt_config* cfg) { return guard([&] { *cfg = C(ctx)->cfg; }); }
int wpt_ctx_render_exact(wpt_ctx* ctx, uint32_t spp) {
  return guard([&] { Context* c = C(ctx); if (c->cfg.render_type == WPT_PNEE) c->build_photons(); c->render_exact(spp); });
}
int64_t wpt_ctx_render_adaptive(wpt_ctx* ctx, uint64_t budget) {
  int64_t used = -1;
  guard([&] { Context* c = C(ctx); if (c->cfg.render_type == WPT_PNEE) c->build_photons(); used = (int64_t)c->render_adaptive(budget); });
  return used;
}
int wpt_ctx_render_random(wpt_ctx* ctx, uint64_t ticks) {
  return guard([&] { Context* c = C(ctx); if (c->cfg.render_type == WPT_PNEE) c->build_photons(); c->render_random(ticks); });
}
int wpt_ctx_set_exchange_callback(wpt_ctx* ctx, void (*cb)(void*), void* user) {
  return guard([&] { Context* c = C(ctx); if (cb) c->exchange_hook = [cb, user] { cb(user); }; else c->exchange_hook = nullptr; });
}
int wpt_ctx_set_reduce_callback(wpt_ctx* ctx, void (*cb)(void*, void*, uint64_t), void* user) {
  return guard([&] { Context* c = C(ctx); if (cb) c->reduce_hook = [cb, user](uint32_t* p, uint64_t n) { cb(user, p, n); }; else c->reduce_hook = nullptr; });
}
int wpt_ctx_build_photons(wpt_ctx* ctx) { return guard([&] { C(ctx)->build_photons(); }); }
// ---- native multi-GPU plane (dist_nccl.cpp)
int wpt_nccl_unique_id(uint8_t out[128]) { return guard([&] { if (!out) throw std::runtime_error("null id buffer"); nccl_unique_id(out); }); }
int wpt_ctx_attach_nccl(wpt_ctx* ctx, const uint8_t id[128], uint32_t rank, uint32_t world) {
  return guard([&] { if (!id && world > 1) throw std::runtime_error("null NCCL id"); C(ctx)->attach_nccl(id, nullptr, rank, world); });
}
int wpt_ctx_attach_nccl_comm(wpt_ctx* ctx, void* comm, uint32_t rank, uint32_t world) {
  return guard([&] { if (!comm && world > 1) throw std::runtime_error("null NCCL communicator"); C(ctx)->attach_nccl(nullptr, comm, rank, world); });
}
int wpt_ctx_detach_nccl(wpt_ctx* ctx) { return guard([&] { Context* c = C(ctx); c->detach_nccl(); c->cfg.rank = 0; c->cfg.world = 1; c->slots = 0; }); }
int wpt_ctx_gather_frame(wpt_ctx* ctx) { return guard([&] { C(ctx)->exchange_native(); }); }
int wpt_ctx_synchronize(wpt_ctx* ctx) { return guard([&] { Context* c = C(ctx); c->require_device(); WPT_CUDA(cudaStreamSynchronize(c->stream)); }); }
int wpt_ctx_stats(wpt_ctx* ctx, uint64_t out[8]) { return guard([&] { C(ctx)->stats(out); }); }

int wpt_ctx_primary_probe(wpt_ctx* ctx, int32_t* ids, uint32_t* visits, float* dist) {
  return guard([&] {
    Context* c = C(ctx);
    c->require_device();
    size_t n = (size_t)c->W * c->H;
    DevBuf<int32_t> d_ids; DevBuf<uint32_t> d_vis; DevBuf<float> d_dist;
    d_ids.alloc(n); d_vis.alloc(n); d_dist.alloc(n);
    launch_primary_probe(c->params(c->cfg.render_type), d_ids.p, d_vis.p, d_dist.p, c->stream);
    if (ids) WPT_CUDA(cudaMemcpyAsync(ids, d_ids.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (visits) WPT_CUDA(cudaMemcpyAsync(visits, d_vis.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (dist) WPT_CUDA(cudaMemcpyAsync(dist, d_dist.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    WPT_CUDA(cudaStreamSynchronize(c->stream));
  });
}
int wpt_ctx_accum(wpt_ctx* ctx, float* rgb, uint32_t* counts) {
  return guard([&] {
    Context* c = C(ctx);
    c->require_device();
    size_t n = (size_t)c->W * c->H;
    std::vector<float4> h(n);
    WPT_CUDA(cudaMemcpyAsync(h.data(), c->d_accum.p, n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    WPT_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; i++) {
      if (rgb) { rgb[i * 3] = h[i].x; rgb[i * 3 + 1] = h[i].y; rgb[i * 3 + 2] = h[i].z; }
      if (counts) std::memcpy(&counts[i], &h[i].w, 4);
    }
  });
}
int wpt_ctx_trace_rays(wpt_ctx* ctx, const float* o, const float* d, uint64_t n, int32_t* ids, float* dist, uint32_t* visits, float* normals) {
  return guard([&] {
    Context* c = C(ctx);
    c->require_device();
    if (!n) return;
    DevBuf<float> d_o, d_d, d_dist, d_n; DevBuf<int32_t> d_ids; DevBuf<uint32_t> d_vis;
    d_o.alloc(n * 3); d_d.alloc(n * 3); d_dist.alloc(n); d_ids.alloc(n); d_vis.alloc(n);
    if (normals) d_n.alloc(n * 3);
    WPT_CUDA(cudaMemcpyAsync(d_o.p, o, n * 12, cudaMemcpyHostToDevice, c->stream));
    WPT_CUDA(cudaMemcpyAsync(d_d.p, d, n * 12, cudaMemcpyHostToDevice, c->stream));
    launch_trace_batch(c->params(c->cfg.render_type), d_o.p, d_d.p, n, d_ids.p, d_dist.p, d_vis.p, d_n.p, c->stream);
    WPT_CUDA(cudaMemcpyAsync(ids, d_ids.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    WPT_CUDA(cudaMemcpyAsync(dist, d_dist.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    WPT_CUDA(cudaMemcpyAsync(visits, d_vis.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (normals) WPT_CUDA(cudaMemcpyAsync(normals, d_n.p, n * 12, cudaMemcpyDeviceToHost, c->stream));
    WPT_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int wpt_ctx_scene_info(wpt_ctx* ctx, uint64_t info[8]) {
  return guard([&] {
    const HostScene& s = C(ctx)->scene;
    info[0] = s.shapes.size(); info[1] = s.num_inf; info[2] = s.lights.size(); info[3] = s.bvh2.size();
    info[4] = s.depth2; info[5] = s.bvh4.size(); info[6] = s.bvh_kind; info[7] = s.depth4;
  });
}
int wpt_ctx_bvh2(wpt_ctx* ctx, float* bounds6, uint32_t* lf, uint32_t* cnt) {
  return guard([&] {
    const HostScene& s = C(ctx)->scene;
    for (size_t i = 0; i < s.bvh2.size(); i++) {
      for (int k = 0; k < 3; k++) { bounds6[i * 6 + k] = s.bvh2[i].box.lo[k]; bounds6[i * 6 + 3 + k] = s.bvh2[i].box.hi[k]; }
      lf[i] = s.bvh2[i].left_first; cnt[i] = s.bvh2[i].count;
    }
  });
}
int wpt_ctx_bvh4(wpt_ctx* ctx, float* bounds24, int32_t* children4, uint32_t* nc) {
  return guard([&] {
    const HostScene& s = C(ctx)->scene;
    for (size_t i = 0; i < s.bvh4.size(); i++) {
      for (int c = 0; c < 4; c++) {
        for (int k = 0; k < 3; k++) { bounds24[i * 24 + c * 6 + k] = s.bvh4[i].child[c].lo[k]; bounds24[i * 24 + c * 6 + 3 + k] = s.bvh4[i].child[c].hi[k]; }
        children4[i * 4 + c] = s.bvh4[i].children[c];
      }
      nc[i] = s.bvh4[i].num_children;
    }
  });
}
int wpt_ctx_shape_order(wpt_ctx* ctx, int32_t* src, int32_t* type) {
  return guard([&] { const HostScene& s = C(ctx)->scene; for (size_t i = 0; i < s.shapes.size(); i++) { src[i] = s.shapes[i].source; type[i] = (int32_t)s.shapes[i].type; } });
}
int wpt_ctx_lights(wpt_ctx* ctx, uint32_t* out) {
  return guard([&] { const HostScene& s = C(ctx)->scene; for (size_t i = 0; i < s.lights.size(); i++) out[i] = s.lights[i]; });
}

int64_t wpt_ctx_photon_count(wpt_ctx* ctx, uint64_t* shots) {
  int64_t n = -1;
  guard([&] { Context* c = C(ctx); if (!c->photons_ready) throw std::runtime_error("photon tree not built"); if (shots) *shots = c->photon_shots; n = (int64_t)c->photon_count; });
  return n;
}
int wpt_ctx_photon_list(wpt_ctx* ctx, uint32_t* light, float* loc3, float* weight) {
  return guard([&] {
    Context* c = C(ctx);
    if (!c->photons_ready) throw std::runtime_error("photon tree not built");
    c->photon_list_host();
    if (light) std::copy(c->ph_light.begin(), c->ph_light.end(), light);
    if (loc3) std::copy(c->ph_loc.begin(), c->ph_loc.end(), loc3);
    if (weight) std::copy(c->ph_w.begin(), c->ph_w.end(), weight);
  });
}
int64_t wpt_ctx_photon_tree(wpt_ctx* ctx, uint32_t* meta, float* cum, float* bins) {
  int64_t n = -1;
  guard([&] {
    Context* c = C(ctx);
    if (!c->photons_ready) throw std::runtime_error("photon tree not built");
    c->photon_tree_host();
    if (meta) std::copy(c->pt_meta.begin(), c->pt_meta.end(), meta);
    if (cum) std::copy(c->pt_cum.begin(), c->pt_cum.end(), cum);
    if (bins) std::copy(c->pt_bins.begin(), c->pt_bins.end(), bins);
    n = (int64_t)(c->pt_meta.size() / 3);
  });
  return n;
}
int wpt_ctx_photon_sample(wpt_ctx* ctx, const float* pts3, const uint32_t* seeds, uint64_t n, uint32_t* light, float* pdf) {
  return guard([&] { C(ctx)->photon_sample_batch(pts3, seeds, n, light, pdf); });
}
int wpt_ctx_error_map(wpt_ctx* ctx, float* mse, float stats3[3]) { return guard([&] { C(ctx)->error_map(mse, stats3); }); }
int wpt_ctx_round_spp(wpt_ctx* ctx, uint32_t* spp) { return guard([&] { C(ctx)->round_spp(spp); }); }

int wpt_ctx_set_stream(wpt_ctx* ctx, uint64_t cuda_stream) {
  return guard([&] {
    Context* c = C(ctx);
    c->require_device();
    WPT_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>((uintptr_t)cuda_stream) : c->own_stream;
  });
}
int64_t wpt_ctx_upload_scene(wpt_ctx* ctx) {
  int64_t n = -1;
  guard([&] { n = C(ctx)->reupload_scene(); });
  return n;
}
int wpt_ctx_profile(wpt_ctx* ctx, int enable) { return guard([&] { C(ctx)->set_profiling(enable != 0); }); }
int wpt_ctx_profile_read(wpt_ctx* ctx, double out[8]) { return guard([&] { C(ctx)->profile_read(out); }); }
int wpt_ctx_profile_read_rounds(wpt_ctx* ctx, double out[8]) { return guard([&] { C(ctx)->profile_read_rounds(out); }); }

int wpt_ctx_device_buffers(wpt_ctx* ctx, uint64_t ptrs[8], uint64_t sizes[8]) {
  return guard([&] {
    Context* c = C(ctx);
    c->require_device();
    size_t n = (size_t)c->W * c->H;
    std::memset(ptrs, 0, 8 * sizeof(uint64_t)); std::memset(sizes, 0, 8 * sizeof(uint64_t));
    ptrs[0] = (uint64_t)(uintptr_t)c->d_accum.p; sizes[0] = n * sizeof(float4);
    ptrs[1] = (uint64_t)(uintptr_t)c->d_rgba.p; sizes[1] = n * 4;
    ptrs[2] = (uint64_t)(uintptr_t)c->d_sampling.p; sizes[2] = n * 4;
    ptrs[3] = (uint64_t)(uintptr_t)c->stream; sizes[3] = 0;
  });
}
int wpt_ctx_mark_accum_dirty(wpt_ctx* ctx) { return guard([&] { C(ctx)->rgba_stale = true; }); }

int64_t wpt_parse_obj(const char* text, uint64_t len, int scale, float* out, uint64_t cap) {
  int64_t n = -1;
  guard([&] {
    std::vector<float> v = parse_obj_text(text, (size_t)len, scale != 0);
    for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
    n = (int64_t)v.size();
  });
  return n;
}
int64_t wpt_ctx_load_obj(wpt_ctx* ctx, uint32_t id, const char* path, int scale) {
  int64_t n = -1;
  guard([&] {
    Context* c = C(ctx);
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    std::stringstream ss; ss << f.rdbuf();
    std::string text = ss.str();
    std::vector<float> v = parse_obj_text(text.data(), text.size(), scale != 0);
    // the same three steps the worker performs (worker.ts:171-179)
    c->mesh_tris.erase(id);
    c->mesh_preload[id] = v;
    n = (int64_t)(v.size() / 3);
  });
  if (n >= 0 && wpt_ctx_notify_mesh_loaded(ctx, id) < 0) return -1;
  return n;
}

// ------------------------------------------------------------------ global instance
void wpt_init(uint32_t w, uint32_t h, uint32_t sid, float cx, float cy, float cz, float rx, float ry) {
  if (g_ctx) { fail("Cannot init again"); return; }   // wasm_interface.rs:74-76
  g_ctx = reinterpret_cast<Context*>(wpt_ctx_create(-1, w, h, sid, cx, cy, cz, rx, ry));
}
const uint8_t* wpt_results(uint32_t show) { return wpt_ctx_results(reinterpret_cast<wpt_ctx*>(g_ctx), show); }
void wpt_reset(void) { wpt_ctx_reset(reinterpret_cast<wpt_ctx*>(g_ctx)); }
void wpt_update_scene(uint32_t id) { wpt_ctx_update_scene(reinterpret_cast<wpt_ctx*>(g_ctx), id); }
void wpt_update_settings(uint32_t lt, uint32_t rt, uint32_t la, uint32_t ra, uint32_t dbg) { wpt_ctx_update_settings(reinterpret_cast<wpt_ctx*>(g_ctx), lt, rt, la, ra, dbg); }
void wpt_update_viewport(uint32_t w, uint32_t h) { wpt_ctx_update_viewport(reinterpret_cast<wpt_ctx*>(g_ctx), w, h); }
void wpt_update_camera(float x, float y, float z, float rx, float ry) { wpt_ctx_update_camera(reinterpret_cast<wpt_ctx*>(g_ctx), x, y, z, rx, ry); }
void wpt_allocate_mesh(uint32_t id, uint32_t nv) { wpt_ctx_allocate_mesh(reinterpret_cast<wpt_ctx*>(g_ctx), id, nv); }
float* wpt_mesh_vertices(uint32_t id) { return wpt_ctx_mesh_vertices(reinterpret_cast<wpt_ctx*>(g_ctx), id); }
int wpt_notify_mesh_loaded(uint32_t id) { return wpt_ctx_notify_mesh_loaded(reinterpret_cast<wpt_ctx*>(g_ctx), id); }
uint8_t* wpt_allocate_texture(uint32_t id, uint32_t w, uint32_t h) { return wpt_ctx_allocate_texture(reinterpret_cast<wpt_ctx*>(g_ctx), id, w, h); }
int wpt_notify_texture_loaded(uint32_t id) { return wpt_ctx_notify_texture_loaded(reinterpret_cast<wpt_ctx*>(g_ctx), id); }
void wpt_compute(uint64_t n) { wpt_ctx_compute(reinterpret_cast<wpt_ctx*>(g_ctx), n); }

}  // extern "C"
