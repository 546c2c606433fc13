// Session logic behind the C ABI: scene (re)build + upload, render targets, the wavefront
// driver loop. No CPU rendering path exists here: without a CUDA device every entry point
// fails.
#include "context.h"
#include <algorithm>
#include <string>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace wpt {

void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw CudaError(std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
}

Context::Context(int dev, uint32_t w, uint32_t h, uint32_t sid, const float cam5[5]) {
  if (dev == -2) {   // host-only: scenes, BVHs and OBJ parsing can be inspected without a GPU
    has_device = false; device = -2;
    W = w; H = h;
    std::memcpy(cam, cam5, sizeof cam);
    wpt_default_config(&cfg);
    select_scene(sid);
    return;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) throw CudaError("no CUDA device available (libwpt has no CPU fallback)");
  if (dev < 0) WPT_CUDA(cudaGetDevice(&dev));
  device = dev;
  WPT_CUDA(cudaSetDevice(device));
  WPT_CUDA(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking));
  stream = own_stream;
  W = w; H = h;
  std::memcpy(cam, cam5, sizeof cam);
  wpt_default_config(&cfg);
  WPT_CUDA(cudaMallocHost((void**)&h_ring, 64 * sizeof(uint32_t)));
  WPT_CUDA(cudaMallocHost((void**)&h_counters, 16 * sizeof(unsigned long long)));
  try {
    w_shadow_n.alloc(2); w_ring.alloc(64); w_counters.alloc(16); w_work.alloc(4);
    WPT_CUDA(cudaMemsetAsync(w_counters.p, 0, 16 * sizeof(unsigned long long), stream));
    alloc_targets();
    select_scene(sid);
  } catch (...) {   // the destructor does not run for a half-built object: release the stream and the pinned buffers here
    if (own_stream) { cudaStreamSynchronize(own_stream); cudaStreamDestroy(own_stream); }
    if (h_scene_blob) cudaFreeHost(h_scene_blob);
    if (h_rgba) cudaFreeHost(h_rgba);
    if (h_sampling) cudaFreeHost(h_sampling);
    if (h_ring) cudaFreeHost(h_ring);
    if (h_counters) cudaFreeHost(h_counters);
    throw;
  }
}

Context::~Context() {
  if (!has_device) return;
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  try { detach_nccl(); } catch (...) {}
  for (auto& e : ev_pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (auto& e : ev_free) cudaEventDestroy(e);
  if (own_stream) cudaStreamDestroy(own_stream);
  if (h_scene_blob) cudaFreeHost(h_scene_blob);
  if (h_rgba) cudaFreeHost(h_rgba);
  if (h_sampling) cudaFreeHost(h_sampling);
  if (h_ring) cudaFreeHost(h_ring);
  if (h_counters) cudaFreeHost(h_counters);
}

void Context::alloc_targets() {
  if (!has_device) return;
  size_t n = (size_t)W * H;
  d_accum.release(); d_rgba.release(); d_sampling.release();
  d_accum.alloc(n); d_rgba.alloc(n * 4); d_sampling.alloc(n * 4);
  if (h_rgba) cudaFreeHost(h_rgba);
  if (h_sampling) cudaFreeHost(h_sampling);
  h_rgba = h_sampling = nullptr;
  WPT_CUDA(cudaMallocHost((void**)&h_rgba, n * 4 + 4));
  WPT_CUDA(cudaMallocHost((void**)&h_sampling, n * 4 + 4));
  slots = 0; slot_region[2] = 0;
  clear_targets();
}

void Context::clear_targets() {   // RenderTarget::clear / new, render_target.rs:31-53
  if (!has_device) return;
  size_t n = (size_t)W * H;
  WPT_CUDA(cudaMemsetAsync(d_accum.p, 0, n * sizeof(float4), stream));
  launch_fill_region_rgba(d_sampling.p, W, 0, 0, W, H, 0xFF000000u, stream);   // black, A = 255
  rgba_stale = true;
}

void Context::reset() {   // wasm_interface.rs:137-148
  if (!has_device) return;
  clear_targets();
  clear_strategies();
  unsigned long long now[4];
  read_counters(now);
  for (int i = 0; i < 4; i++) life[i] += now[i];
  WPT_CUDA(cudaMemsetAsync(w_counters.p, 0, 16 * sizeof(unsigned long long), stream));
  photon_rays = photon_visits = 0;
  iterations = 0; launches = 1;   // the sampling-view fill of clear_targets() above
  photons_shot_total = photons_stored_total = 0;
}

void Context::select_scene(uint32_t id) {
  std::vector<HostShape> shapes; std::vector<HostMaterial> mats;
  float bg[3] = {0.0f, 0.0f, 0.0f};   // Color3::BLACK for both reference scenes (scenes.rs:51,110)
  if (id == WPT_SCENE_MUSEUM) scene_museum(shapes, mats);
  else if (id == WPT_SCENE_BUNNY) {
    auto it = mesh_tris.find(1);   // MESH_BUNNY_HIGH, scenes.rs:12
    scene_bunny(it == mesh_tris.end() ? nullptr : &it->second, shapes, mats);
  } else if (id == WPT_SCENE_EXT_WHITTED) {   // extension: not a reference scene id
    auto it = textures.find(0);
    scene_whitted(it != textures.end(), shapes, mats, bg);
  } else throw std::runtime_error("Invalid scene");
  HostScene ns;
  build_scene(ns, std::move(shapes), std::move(mats), cfg.bvh_kind);
  ns.bg[0] = bg[0]; ns.bg[1] = bg[1]; ns.bg[2] = bg[2];
  if (ns.depth2 + 2 > 64 || ns.depth4 * 3 + 4 > 64) throw std::runtime_error("BVH too deep for the device traversal stack");
  for (const HostMaterial& m : ns.mats)   // validate before anything is replaced (the upload re-allocates the device buffers)
    if (m.kind == MAT_DIFFUSE_TEX) {
      auto dm = tex_dims.find(m.tex);
      if (m.tex >= WPT_MAX_TEXTURES || textures.find(m.tex) == textures.end() || dm == tex_dims.end() || !dm->second.first || !dm->second.second)
        throw std::runtime_error("textured material without a loaded texture");
    }
  // the session keeps its old scene unless the new one is completely uploaded
  HostScene old = std::move(scene);
  scene = std::move(ns);
  try { upload_scene(); }
  catch (...) {
    scene = std::move(old);
    try { if (!scene.shapes.empty()) upload_scene(); } catch (...) {}
    throw;
  }
  scene_id = id;
  photons_ready = false; photon_shots = photon_count = 0;
}

void Context::upload_scene() {
  scene_version++;
  if (!has_device) return;
  std::vector<DNode2> n2; std::vector<DNode4> n4; std::vector<DShape> shp; std::vector<DMaterial> mats; std::vector<DLight> lights;
  flatten_scene(scene, n2, n4, shp, mats, lights);
  WPT_CUDA(cudaStreamSynchronize(stream));   // nothing may still read the old buffers
  // keep one pinned host blob of the flattened scene: [nodes2 | nodes4 | shapes | mats | lights]
  const void* src[5] = {n2.data(), n4.data(), shp.data(), mats.data(), lights.data()};
  size_t len[5] = {n2.size() * sizeof(DNode2), n4.size() * sizeof(DNode4), shp.size() * sizeof(DShape), mats.size() * sizeof(DMaterial), lights.size() * sizeof(DLight)};
  size_t total = 0;
  for (int i = 0; i < 5; i++) { blob_off[i] = total; blob_len[i] = len[i]; total += (len[i] + 255) & ~(size_t)255; }
  if (h_scene_blob) { cudaFreeHost(h_scene_blob); h_scene_blob = nullptr; }
  WPT_CUDA(cudaMallocHost(&h_scene_blob, total ? total : 256));
  h_scene_bytes = total;
  for (int i = 0; i < 5; i++) if (len[i]) std::memcpy((char*)h_scene_blob + blob_off[i], src[i], len[i]);
  d_nodes2.alloc(n2.size()); d_nodes4.alloc(n4.size()); d_shapes.alloc(shp.size()); d_mats.alloc(mats.size()); d_lights.alloc(lights.size());
  for (uint32_t t = 0; t < WPT_MAX_TEXTURES; t++) {   // textures referenced by textured-diffuse materials (extension scene only)
    d_tex_w[t] = d_tex_h[t] = 0;
    auto it = textures.find(t); auto dm = tex_dims.find(t);
    if (it == textures.end() || dm == tex_dims.end() || !dm->second.first || !dm->second.second) continue;
    size_t bytes = (size_t)dm->second.first * dm->second.second * 3;
    d_tex[t].alloc(bytes);
    WPT_CUDA(cudaMemcpyAsync(d_tex[t].p, it->second.data(), bytes, cudaMemcpyHostToDevice, stream));
    d_tex_w[t] = dm->second.first; d_tex_h[t] = dm->second.second;
  }
  for (const HostMaterial& m : scene.mats)
    if (m.kind == MAT_DIFFUSE_TEX && (m.tex >= WPT_MAX_TEXTURES || !d_tex_w[m.tex])) throw std::runtime_error("textured material without a loaded texture");
  reupload_scene();
  WPT_CUDA(cudaStreamSynchronize(stream));
}

int64_t Context::reupload_scene() {
  require_device();
  void* dst[5] = {d_nodes2.p, d_nodes4.p, d_shapes.p, d_mats.p, d_lights.p};
  int64_t bytes = 0;
  for (int i = 0; i < 5; i++)
    if (blob_len[i]) { WPT_CUDA(cudaMemcpyAsync(dst[i], (char*)h_scene_blob + blob_off[i], blob_len[i], cudaMemcpyHostToDevice, stream)); bytes += (int64_t)blob_len[i]; }
  return bytes;
}

void Context::read_counters(unsigned long long out[4]) {
  WPT_CUDA(cudaMemcpyAsync(h_counters, w_counters.p, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  WPT_CUDA(cudaStreamSynchronize(stream));
  for (int i = 0; i < 4; i++) out[i] = h_counters[i];
}

cudaEvent_t Context::ev_get() {
  if (!ev_free.empty()) { cudaEvent_t e = ev_free.back(); ev_free.pop_back(); return e; }
  cudaEvent_t e;
  WPT_CUDA(cudaEventCreate(&e));
  return e;
}
void Context::ev_harvest() {   // call after a stream synchronize
  for (auto& p : ev_pending) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { prof_ms[p.kind] += ms; prof_n[p.kind]++; }
    ev_free.push_back(p.a); ev_free.push_back(p.b);
  }
  ev_pending.clear();
}
void Context::ev_mark(int kind, bool begin) {
  if (!profiling) return;
  if (begin) { ev_open[kind] = ev_get(); WPT_CUDA(cudaEventRecord(ev_open[kind], stream)); }
  else if (ev_open[kind]) { cudaEvent_t b = ev_get(); WPT_CUDA(cudaEventRecord(b, stream)); ev_pending.push_back(EvPair{ev_open[kind], b, kind}); ev_open[kind] = nullptr; }
}
void Context::set_profiling(bool on) {
  require_device();
  WPT_CUDA(cudaStreamSynchronize(stream));
  ev_harvest();
  profiling = on;
  for (int i = 0; i < 6; i++) { prof_ms[i] = 0; prof_n[i] = 0; ev_open[i] = nullptr; }
  adaptive_rounds = 0;
  read_counters(prof_base);
  for (int i = 0; i < 4; i++) prof_base[i] += life[i];
}
void Context::profile_read_rounds(double out[8]) {   // rounds, error-map ms, render ms, exchange ms, photon warm-up ms (wall), collectives
  require_device();
  WPT_CUDA(cudaStreamSynchronize(stream));
  ev_harvest();
  out[0] = (double)adaptive_rounds; out[1] = prof_ms[2]; out[2] = prof_ms[3]; out[3] = prof_ms[4]; out[4] = photon_build_ms; out[5] = (double)collectives; out[6] = out[7] = 0;
}
void Context::profile_read(double out[8]) {
  require_device();
  WPT_CUDA(cudaStreamSynchronize(stream));
  ev_harvest();
  unsigned long long now[4];
  read_counters(now);
  for (int i = 0; i < 4; i++) now[i] += life[i];
  out[0] = prof_ms[0]; out[1] = (double)prof_n[0]; out[2] = prof_ms[1]; out[3] = (double)prof_n[1];
  out[4] = (double)(now[3] - prof_base[3]); out[5] = (double)(now[0] - prof_base[0]); out[6] = (double)(now[1] - prof_base[1]); out[7] = 0;
}

RenderParams Context::params(uint32_t render_type) const {
  RenderParams rp{};
  rp.scene.nodes2 = d_nodes2.p; rp.scene.nodes4 = d_nodes4.p; rp.scene.shapes = d_shapes.p; rp.scene.mats = d_mats.p; rp.scene.lights = d_lights.p;
  rp.scene.num_inf = scene.num_inf; rp.scene.num_shapes = (uint32_t)scene.shapes.size(); rp.scene.num_lights = (uint32_t)scene.lights.size();
  rp.scene.bvh_kind = scene.bvh_kind;
  rp.scene.bg_r = scene.bg[0]; rp.scene.bg_g = scene.bg[1]; rp.scene.bg_b = scene.bg[2];
  for (uint32_t t = 0; t < WPT_MAX_TEXTURES; t++) { rp.scene.tex[t].rgb = d_tex[t].p; rp.scene.tex[t].width = d_tex_w[t]; rp.scene.tex[t].height = d_tex_h[t]; }
  rp.cam.ox = cam[0]; rp.cam.oy = cam[1]; rp.cam.oz = cam[2];
  // sin/cos of the camera angles are evaluated once on the host (vec3.rs:103-104,116-117)
  rp.cam.cx = std::cos(cam[3]); rp.cam.sx = std::sin(cam[3]); rp.cam.cy = std::cos(cam[4]); rp.cam.sy = std::sin(cam[4]);
  float fw = (float)W, fh = (float)H;
  rp.cam.w_inv = 1.0f / fw; rp.cam.h_inv = 1.0f / fh; rp.cam.ar = fw / fh;   // tracer.rs:168-172
  rp.photons.child_base = p_child_base.p; rp.photons.cum = p_cum.p; rp.photons.nbr = p_nbr.p;
  rp.photons.num_lights = (uint32_t)scene.lights.size(); rp.photons.num_nodes = p_nodes;
  rp.W = W; rp.H = H;
  rp.render_type = render_type; rp.light_debug = cfg.light_debug; rp.base_seed = cfg.base_seed;
  return rp;
}

void Context::region(uint32_t* rx, uint32_t* ry, uint32_t* rw, uint32_t* rh) const {
  *rx = cfg.region_x; *ry = cfg.region_y;
  *rw = cfg.region_w ? cfg.region_w : W;
  *rh = cfg.region_h ? cfg.region_h : H;
  if (*rx + *rw > W || *ry + *rh > H) throw std::runtime_error("region outside the viewport");
}

// One slot per pixel of this session's rows of the region: the 4-row bands b with b % world == rank (wpt_types.h)
void Context::ensure_slots(uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh) {
  require_device();
  uint32_t world = cfg.world ? cfg.world : 1, rank = cfg.rank;
  uint32_t key[6] = {rx, ry, rw, rh, rank, world};
  // the slot order depends on what the camera sees (launch_order_tiles): rebuilt when the camera, the scene or the BVH kind change
  // (off unless WPT_TILE_ORDER=1: measured -1.7 % on the bunny frame, +9 % on the indoor museum — gpurun_out/r2h_ab.log, r2h_zone.log)
  static const bool order_tiles = std::getenv("WPT_TILE_ORDER") && std::atoi(std::getenv("WPT_TILE_ORDER")) != 0;
  const uint64_t view = order_tiles ? view_key() : 0;
  if (slots && !std::memcmp(key, slot_region, sizeof key) && view == slot_view) return;
  uint32_t rows = band_rows(rh, rank, world);
  uint32_t n = rows * rw;
  s_pixel.alloc(n); s_spp.alloc(n);
  if (use_wavefront()) ensure_wavefront_state(n);   // ~140 B per slot that the persistent kernels never touch
  launch_fill_pixels(s_pixel.p, W, rx, ry, rw, rh, rank, world, stream);
  const uint32_t ntiles = ((rw & ~7u) * (rows & ~3u)) >> 5;
  if (order_tiles && ntiles > 1) {
    s_pixel_alt.alloc(n); t_key.alloc(ntiles); t_val.alloc(ntiles); t_key2.alloc(ntiles); t_val2.alloc(ntiles);
    const size_t tb = tile_sort_bytes(ntiles);
    t_tmp.alloc(tb);
    launch_order_tiles(params(cfg.render_type), s_pixel.p, n, ntiles, t_key.p, t_val.p, t_key2.p, t_val2.p, t_tmp.p, tb, s_pixel_alt.p, stream);
    std::swap(s_pixel.p, s_pixel_alt.p); std::swap(s_pixel.n, s_pixel_alt.n);
    launches += 3;
  }
  slots = n;
  slot_view = view;
  std::memcpy(slot_region, key, sizeof key);
}
uint64_t Context::view_key() const {   // FNV-1a over the camera, the scene identity and the BVH kind
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
  mix(cam, sizeof cam); mix(&scene_id, sizeof scene_id); mix(&scene_version, sizeof scene_version); mix(&cfg.bvh_kind, sizeof cfg.bvh_kind);
  return h | 1ull;
}

void Context::ensure_wavefront_state(uint32_t n) {
  s_ray_o.alloc(n); s_ray_d.alloc(n); s_col.alloc(n); s_sh_o.alloc(n); s_sh_d.alloc(n); s_sh_c.alloc(n); s_tail.alloc(n);
  s_misc.alloc(n); s_hit.alloc(n);
  w_shadow_q0.alloc(n); w_shadow_q1.alloc(n);
}

PathState Context::path_state() {
  PathState st{};
  st.ray_o = s_ray_o.p; st.ray_d = s_ray_d.p; st.col = s_col.p; st.misc = s_misc.p; st.hit = s_hit.p;
  st.sh_o = s_sh_o.p; st.sh_d = s_sh_d.p; st.sh_c = s_sh_c.p; st.tail = s_tail.p; st.pixel = s_pixel.p; st.n = slots;
  return st;
}
WaveBuffers Context::wave_buffers() {
  WaveBuffers wb{};
  wb.shadow_q[0] = w_shadow_q0.p; wb.shadow_q[1] = w_shadow_q1.p; wb.shadow_n = w_shadow_n.p; wb.active_ring = w_ring.p;
  wb.counters = w_counters.p; wb.accum = d_accum.p;
  return wb;
}

// The wavefront loop: shade(0) generates the first camera rays, then trace / shade alternate
// until no slot is live. The live count is read back every `poll` iterations.
void Context::run_wavefront(uint32_t render_type, const uint32_t* d_spp_per_slot, uint32_t uniform_spp) {
  require_device();
  if (!slots) return;
  if (render_type == WPT_PNEE && !photons_ready) throw std::runtime_error("photon tree not built");
  ensure_wavefront_state(slots);   // the engine may have been selected after the slots were laid out
  RenderParams rp = params(render_type);
  PathState st = path_state();
  WaveBuffers wb = wave_buffers();
  // contract B10: this run is one segment — its samples are summed from +0 in d_seg_acc, then added
  d_seg_acc.alloc((size_t)W * H);
  launch_clear_pixels(d_seg_acc.p, s_pixel.p, slots, stream);
  wb.accum = d_seg_acc.p;
  launch_setup_slots(st, d_spp_per_slot, uniform_spp, d_accum.p, stream);
  WPT_CUDA(cudaMemsetAsync(w_shadow_n.p, 0, 2 * sizeof(uint32_t), stream));
  WPT_CUDA(cudaMemsetAsync(w_ring.p, 0, 64 * sizeof(uint32_t), stream));
  launches += 1;
  const int grid = device_sm_count() * 8;
  const uint32_t poll = 8;
  uint32_t iter = 0;
  launch_shade(rp, st, wb, iter, grid, stream);
  launches += 1;
  for (;;) {
    for (uint32_t k = 0; k < poll; k++) {
      iter++;
      if (profiling) {
        EvPair t{ev_get(), ev_get(), 0}, sh{ev_get(), ev_get(), 1};
        WPT_CUDA(cudaEventRecord(t.a, stream));
        launch_trace(rp, st, wb, iter, grid, stream);
        WPT_CUDA(cudaEventRecord(t.b, stream));
        WPT_CUDA(cudaEventRecord(sh.a, stream));
        launch_shade(rp, st, wb, iter, grid, stream);
        WPT_CUDA(cudaEventRecord(sh.b, stream));
        ev_pending.push_back(t); ev_pending.push_back(sh);
      } else {
        launch_trace(rp, st, wb, iter, grid, stream);
        launch_shade(rp, st, wb, iter, grid, stream);
      }
      launches += 2;
    }
    WPT_CUDA(cudaMemcpyAsync(h_ring, w_ring.p, 64 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    WPT_CUDA(cudaStreamSynchronize(stream));
    if (profiling) ev_harvest();
    if (h_ring[iter & 63u] == 0) break;
  }
  launch_add_segment(d_accum.p, s_pixel.p, slots, d_seg_acc.p, stream);
  launches += 2;
  iterations += iter;
  rgba_stale = true;
}

// The persistent engine: one launch renders every requested sample of every slot.
void Context::run_persistent(uint32_t render_type, const uint32_t* d_spp_per_slot, uint32_t uniform_spp) {
  require_device();
  if (!slots) return;
  if (render_type == WPT_PNEE && !photons_ready) throw std::runtime_error("photon tree not built");
  MegaParams P{};
  // kernel variant by scene content: triangles + planes only / the reference's primitives and materials / everything
  uint32_t scene_kind = 0;
  for (const HostShape& sh : scene.shapes) {
    if (sh.type == SH_SPHERE || sh.type == SH_SQUARE) { scene_kind = 2; break; }
    if (sh.type != SH_TRIANGLE && sh.type != SH_PLANE) scene_kind = 1;
  }
  for (const HostMaterial& m : scene.mats) if (m.kind != MAT_DIFFUSE && m.kind != MAT_EMISSIVE) { scene_kind = 2; break; }
  if (std::getenv("WPT_NO_SIMPLE")) scene_kind = 2;
  const bool zones_ok = scene_kind != 0;   // the triangles / planes variants are compiled without the end-zone code (kernels.cu)
  for (int z = 0; z < 3; z++) { P.zone_start[z] = 0xFFFFFFFFu; P.zone_pslot[z] = 0; P.zone_per[z] = 1; P.zone_len[z] = 1; }
  P.zone_samples = 0; uint32_t zone_first = slots;
  P.rp = params(render_type);
  P.accum = d_accum.p; P.pixel = s_pixel.p; P.spp_per_slot = d_spp_per_slot; P.uniform_spp = uniform_spp;
  // contract B10: render_exact cuts a pixel's samples into segments of WPT_SEGMENT_LEN, each summed from +0 by its
  // own slot (so one pixel's samples can run on several lanes); a strategy round (per-slot counts) is one segment.
  const uint32_t seg_len = WPT_SEGMENT_LEN;   // 8 segments of 8 samples cover the 64 samples a strategy round may hold per pixel (the list encoding has 3 segment bits)
  P.seg_len = seg_len;
  if (d_spp_per_slot) {
    // strategy round: per-slot counts (0..33 adaptive, anything for random). The segments of all pixels are listed
    // pixel by pixel on the device; pixels without samples get no slot. At most 8 segments per pixel fit the list
    // encoding (pixel slot << 3 | segment); the strategies cap a round at 33 samples, random rounds are checked.
    if (slots >= (1u << 26)) throw std::runtime_error("viewport too large for the segment list encoding");
    // Short slots: a round is one launch, and with few pixel slots per lane (small frames, many GPUs) it lasts as long as its slowest
    // slot — ~1 ms per segment of 8 samples plus the stragglers (scripts/tail_probe.py). Below four pixel slots per lane the round is
    // cut into slots of 2 samples (WPT_MEGA_LIST_LEN overrides; 0 = segments); they store one colour per sample and
    // k_combine_segment_list forms the segment sums in the same order, so the slot length is a scheduling choice, not part of contract B10.
    static const int env_list = std::getenv("WPT_MEGA_LIST_LEN") ? std::atoi(std::getenv("WPT_MEGA_LIST_LEN")) : -1;
    const uint32_t lanes = (uint32_t)device_sm_count() * 8u * 128u;
    uint32_t list_len = env_list >= 0 ? (uint32_t)std::min(env_list, 64) : (slots < 4u * lanes ? 2u : 0u);
    if (cfg.engine != 0) list_len = 0;
    if (list_len && (uint64_t)slots * ((64u + list_len - 1u) / list_len) * list_len >= (1ull << 31)) list_len = 0;   // per-sample indices carry a flag in bit 31
    const uint32_t per_px = list_len ? (64u + list_len - 1u) / list_len : 8u;   // slots a pixel can have (a round holds at most 64 samples per pixel)
    d_seg_cnt.alloc((size_t)slots + 1); d_seg_off.alloc((size_t)slots + 1); d_seg_list.alloc((size_t)slots * per_px);
    size_t sb = seg_scan_bytes(slots);
    d_scan_tmp.alloc(sb ? sb : 1);
    launch_build_segment_list(d_spp_per_slot, slots, list_len ? list_len : seg_len, d_seg_cnt.p, d_seg_off.p, d_scan_tmp.p, sb, d_seg_list.p, list_len ? 6u : 3u, stream);
    d_seg_buf.alloc(list_len ? (size_t)slots * per_px * list_len : (size_t)slots * 8u);
    P.list_len = list_len;
    P.nseg = 1; P.seg_list = d_seg_list.p; P.nslots_dev = d_seg_off.p + slots; P.nslots = slots; P.seg_buf = d_seg_buf.p;
    launches += 3;
  } else {
    P.nseg = std::max(1u, (uniform_spp + seg_len - 1u) / seg_len);
    // End zones: a launch ends when its last slot ends, and a lane needs 25 - 40 us per ray, i.e. ~1 ms for a segment of 8 samples:
    // when the queue runs dry every lane still holds half a segment on average and the slowest a whole one (scripts/tail_probe.py:
    // the median warp exits 1.0 ms, the last 2.5 ms after the queue is empty). So the last part of the pixel slots is queued in
    // shorter slots, graded (WPT_MEGA_ZONES = "percent:samples,..." in queue order, e.g. "15:4,10:2,5:1"): each zone has to last
    // as long as the slowest slot of the zone before it. Zone slots store one colour per sample and k_combine_segments forms the
    // segment sums in the same order (same bits).
    // Measured (gpurun_out/r2h_zone2.log, r2h_tail_zone2.log): "15:4,10:2,5:1" pulls the warp exits together (99 % of the warps within
    // 0.47 ms of the queue running dry instead of 1.76 ms, the last one after 1.6 instead of 2.6 ms), but every slot fetch stalls its
    // warp on the atomic + pixel + accumulator loads, so the queue itself lasts longer: bunny frame 16.6 -> 16.8 ms (no gain), museum
    // 32.5 -> 31.3 ms. So only the kernel variants for scenes with tori / boxes / the extension carry the zone code, on by default.
    static const char* env_zones_c = std::getenv("WPT_MEGA_ZONES");
    const std::string env_zones = !zones_ok ? "" : (env_zones_c ? env_zones_c : "15:4,10:2,5:1");
    uint32_t zpx[3] = {0, 0, 0}, zlen[3] = {1, 1, 1}; int nz = 0;
    if (uniform_spp > 1 && cfg.engine == 0) {
      const char* c = env_zones.c_str();
      while (*c && nz < 3) {
        char* e = nullptr; long pct = std::strtol(c, &e, 10); if (e == c || *e != ':') break;
        c = e + 1; long len = std::strtol(c, &e, 10); if (e == c) break;
        c = (*e == ',') ? e + 1 : e;
        if (pct > 0 && len > 0) { zpx[nz] = (uint32_t)((uint64_t)slots * (uint32_t)std::min(100l, pct) / 100u) & ~31u; zlen[nz] = (uint32_t)len; nz++; }
      }
    }
    uint32_t zone_px = zpx[0] + zpx[1] + zpx[2];
    if (zone_px > slots) { zone_px = 0; nz = 0; }
    const uint32_t body_px = slots - zone_px;
    uint64_t total = (uint64_t)body_px * P.nseg;
    {
      uint32_t px0 = body_px;
      for (int z = 0; z < 3; z++) {
        P.zone_start[z] = 0xFFFFFFFFu; P.zone_pslot[z] = px0; P.zone_len[z] = zlen[z]; P.zone_per[z] = (uniform_spp + zlen[z] - 1u) / zlen[z];
        if (z < nz && zpx[z]) { P.zone_start[z] = (uint32_t)std::min<uint64_t>(total, 0xFFFFFFFEull); total += (uint64_t)zpx[z] * P.zone_per[z]; px0 += zpx[z]; }
      }
    }
    if (total > 0x7FFFFFFFull) throw std::runtime_error("too many samples per pixel for one render_exact call at this viewport size");
    P.nslots = (uint32_t)total;
    zone_first = body_px;
    P.zone_samples = body_px * P.nseg;
    const uint64_t buf_n = (uint64_t)body_px * P.nseg + (uint64_t)zone_px * uniform_spp;
    if (P.nseg > 1 || zone_px) { d_seg_buf.alloc((size_t)buf_n); P.seg_buf = d_seg_buf.p; }
  }
  P.seg_buf_n = P.seg_buf ? (uint32_t)std::min<size_t>(d_seg_buf.n, 0xFFFFFFFFu) : 0u;
  P.work_counter = w_work.p; P.counters = w_counters.p;
  WPT_CUDA(cudaMemsetAsync(w_work.p, 0, sizeof(uint32_t), stream));
#ifdef MEGA_INSTR
  const size_t dbg_n = 128 + 3 * (size_t)device_sm_count() * 16 * 4;
  d_dbg.alloc(dbg_n); WPT_CUDA(cudaMemsetAsync(d_dbg.p, 0, dbg_n * sizeof(unsigned long long), stream)); P.dbg = d_dbg.p;
  { static const unsigned long long init[4] = {~0ull, ~0ull, 0ull, 0ull}; WPT_CUDA(cudaMemcpyAsync(w_counters.p + 9, init, sizeof init, cudaMemcpyHostToDevice, stream)); }
#endif
  cudaEvent_t a = nullptr, b = nullptr;
  if (profiling) { a = ev_get(); b = ev_get(); WPT_CUDA(cudaEventRecord(a, stream)); }
  // burst thresholds: 20 / 10 for the triangles + planes variant; 12 / 4 for scenes with tori, where parked lanes (torus phase) thin the
  // bursts out (museum 8-spp frame: 20/10 36.4 ms, 12/4 34.1 ms, gpurun_out/r2f_museum.log); 0 = not set
  static const int env_hi = std::getenv("WPT_MEGA_THI") ? std::atoi(std::getenv("WPT_MEGA_THI")) : 0;
  static const int env_lo = std::getenv("WPT_MEGA_TLO") ? std::atoi(std::getenv("WPT_MEGA_TLO")) : 0;
  // blocks per SM (= register budget) of the four kernel variants: triangles/planes BVH2, BVH4, generic BVH2, BVH4
  // (measured: with the f64 torus solver inlined at two call sites the generic variant wanted every warp it could get — museum, 8-spp frame:
  //  4 blocks/SM 85 ms, 8: 72, 12: 65, 16: 59 ms; with the deferred torus phase (one call site, run for many parked lanes at once) it is
  //  8: 35.8, 16: 38.3 ms, gpurun_out/r2f_museum.log;
  //  triangles/planes BVH2 8 or 9, 10 is worse; BVH4 5, or 8 with PNEE)
  static const int env_minb_base[4] = {std::getenv("WPT_MEGA_MINB") ? std::atoi(std::getenv("WPT_MEGA_MINB")) : 8, std::getenv("WPT_MEGA_MINB4") ? std::atoi(std::getenv("WPT_MEGA_MINB4")) : 0,
                                       std::getenv("WPT_MEGA_MINBG") ? std::atoi(std::getenv("WPT_MEGA_MINBG")) : 8, std::getenv("WPT_MEGA_MINBG4") ? std::atoi(std::getenv("WPT_MEGA_MINBG4")) : 16};
  int env_minb[4] = {env_minb_base[0], env_minb_base[1] ? env_minb_base[1] : (render_type == WPT_PNEE ? 8 : 5), env_minb_base[2], env_minb_base[3]};
  static const int env_ti = std::getenv("WPT_MEGA_TINNER") ? std::atoi(std::getenv("WPT_MEGA_TINNER")) : 2;
  static const int env_reps = std::getenv("WPT_MEGA_REPS") ? std::atoi(std::getenv("WPT_MEGA_REPS")) : 4;   // measured: 1 -> 20.0, 2 -> 19.3, 4 -> 18.8, 8 -> 19.5 ms (gpurun_out/sweep10*.log)
  P.t_hi = (uint32_t)(env_hi ? env_hi : 20); P.t_lo = (uint32_t)(env_lo ? env_lo : 10); P.t_inner = (uint32_t)env_ti; P.inner_reps = (uint32_t)std::max(1, env_reps);
  static const int env_ttor = std::getenv("WPT_MEGA_TTORUS") ? std::atoi(std::getenv("WPT_MEGA_TTORUS")) : 12;   // measured: 4 39.1, 8 36.4, 10 35.8, 12 35.9 ms at first; 8 29.8, 10 29.1, 12 28.9, 14 29.1 ms at the end of the round (museum, 8 blocks / SM)
  P.t_torus = (uint32_t)std::max(1, env_ttor);
  static const int env_chunk = std::getenv("WPT_MEGA_CHUNK") ? std::atoi(std::getenv("WPT_MEGA_CHUNK")) : 32;
  P.chunk = (uint32_t)env_chunk;
  P.scene_kind = scene_kind;
  if (P.scene_kind != 0) { if (!env_hi) P.t_hi = 12; if (!env_lo) P.t_lo = 4; }
  {
    // the infinite shapes are planes (only a Plane has no finite box, bvh.rs:376-394), at most two in the reference's
    // scenes: they and the BVH2 root node travel in the kernel parameters (constant bank)
    const DShape* shp = reinterpret_cast<const DShape*>((const char*)h_scene_blob + blob_off[2]);
    for (uint32_t i = 0; i < scene.num_inf && i < 2; i++) P.inf_q1[i] = shp[i].q1;
    const DNode2* n2 = reinterpret_cast<const DNode2*>((const char*)h_scene_blob + blob_off[0]);
    if (blob_len[0] >= sizeof(DNode2)) { P.root_a = n2[0].a; P.root_b = n2[0].b; }
  }
  {
    auto mix32 = [](uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; };
    P.seed_path = mix32((uint32_t)STREAM_PATH ^ cfg.base_seed);
  }
  if (cfg.engine == 4) {
#ifndef WPT_EXPERIMENTAL
    throw std::runtime_error("engine 4 (experimental warp-pool kernel) is not in this build: make EXTRA=-DWPT_EXPERIMENTAL EXPERIMENTAL_SRCS=wpool.cu");
#else
    // experimental warp-pool kernel (wpool.cu): grid = SMs x blocks per SM; every warp owns pool_ctx path contexts of 160 B
    static const int w_minb = std::getenv("WPT_WPOOL_MINB") ? std::atoi(std::getenv("WPT_WPOOL_MINB")) : 8;
    static const int w_ctx = std::getenv("WPT_WPOOL_CTX") ? std::atoi(std::getenv("WPT_WPOOL_CTX")) : 96;
    static const int w_thi = std::getenv("WPT_WPOOL_THI") ? std::atoi(std::getenv("WPT_WPOOL_THI")) : 32;
    static const int w_tlo = std::getenv("WPT_WPOOL_TLO") ? std::atoi(std::getenv("WPT_WPOOL_TLO")) : 16;
    static const int w_tsw = std::getenv("WPT_WPOOL_TSWITCH") ? std::atoi(std::getenv("WPT_WPOOL_TSWITCH")) : 16;
    static const int w_ref = std::getenv("WPT_WPOOL_REFILL") ? std::atoi(std::getenv("WPT_WPOOL_REFILL")) : 8;
    const int minb = w_minb >= 12 ? 12 : w_minb >= 8 ? 8 : 6;
    int grid = device_sm_count() * minb;
    if (!P.nslots_dev) grid = std::min(grid, (int)((P.nslots + 127) / 128));
    P.pool_ctx = (uint32_t)std::min(128, std::max(32, w_ctx));
    d_pool.alloc((size_t)wpool_warps(device_sm_count() * minb) * P.pool_ctx * (wpool_ctx_bytes() / sizeof(float4)));
    P.pool = d_pool.p;
    P.t_hi = (uint32_t)w_thi; P.t_lo = (uint32_t)w_tlo; P.t_switch = (uint32_t)w_tsw; P.t_refill = (uint32_t)std::max(1, w_ref);
    launch_wpool(P, grid, minb, stream);
#endif
  } else launch_mega(P, env_minb, stream);
  if (profiling) { WPT_CUDA(cudaEventRecord(b, stream)); ev_pending.push_back(EvPair{a, b, 0}); }
  WPT_CUDA(cudaGetLastError());
  if (P.seg_list) { launch_combine_segment_list(d_accum.p, s_pixel.p, slots, d_seg_buf.p, d_seg_off.p, d_spp_per_slot, P.list_len, seg_len, stream); launches += 1; }
  else if (P.seg_buf) { launch_combine_segments(d_accum.p, s_pixel.p, slots, d_seg_buf.p, P.nseg, uniform_spp, zone_first, seg_len, stream); launches += 1; }
  launches += 1; iterations += 1;
  rgba_stale = true;
}

// cfg.engine: 0 = persistent path kernel k_mega (default), 1 = multi-kernel wavefront, 4 = experimental warp-pool kernel k_wpool
void Context::run_paths(uint32_t render_type, const uint32_t* d_spp_per_slot, uint32_t uniform_spp) {
  if (use_wavefront()) run_wavefront(render_type, d_spp_per_slot, uniform_spp);
  else run_persistent(render_type, d_spp_per_slot, uniform_spp);
}

void Context::render_exact(uint32_t spp) {
  require_device();
  uint32_t rx, ry, rw, rh;
  region(&rx, &ry, &rw, &rh);
  ensure_slots(rx, ry, rw, rh);
  if (use_wavefront()) {   // the wavefront engine runs the segments of contract B10 one after the other
    for (uint32_t rem = spp; rem > 0;) { uint32_t m = std::min(rem, WPT_SEGMENT_LEN); run_wavefront(cfg.render_type, nullptr, m); rem -= m; }
  } else {
    // one launch per call where the segment-sum buffer allows it (16 B per pixel and segment, capped at 2 GiB and 64
    // segments = 512 samples per pixel): one ramp-up and one tail. The segments of later launches are added after
    // those of earlier ones, i.e. in the same order as in one launch.
    const uint64_t per_seg = (uint64_t)std::max(1u, slots) * sizeof(float4);
    const uint32_t max_seg = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(8, (2ull << 30) / per_seg));
    for (uint32_t rem = spp; rem > 0;) { uint32_t m = std::min(rem, max_seg * WPT_SEGMENT_LEN); run_persistent(cfg.render_type, nullptr, m); rem -= m; }
  }
}

const uint8_t* Context::results(uint32_t show_sampling) {   // wasm_interface.rs:120-134
  require_device();
  size_t n = (size_t)W * H;
  if (show_sampling == 1) {
    WPT_CUDA(cudaMemcpyAsync(h_sampling, d_sampling.p, n * 4, cudaMemcpyDeviceToHost, stream));
    WPT_CUDA(cudaStreamSynchronize(stream));
    return h_sampling;
  }
  if (rgba_stale) {
    launch_resolve_rgba(d_accum.p, d_rgba.p, (uint32_t)n, stream);
    WPT_CUDA(cudaMemcpyAsync(h_rgba, d_rgba.p, n * 4, cudaMemcpyDeviceToHost, stream));
    rgba_stale = false;
  }
  WPT_CUDA(cudaStreamSynchronize(stream));
  return h_rgba;
}

void Context::stats(uint64_t out[8]) {
  require_device();
  WPT_CUDA(cudaMemcpyAsync(h_counters, w_counters.p, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  WPT_CUDA(cudaStreamSynchronize(stream));
  if (std::getenv("WPT_DEBUG_COUNTERS")) {
    std::fprintf(stderr, "wpt counters:");
    for (int i = 0; i < 16; i++) std::fprintf(stderr, " %llu", h_counters[i]);
    std::fprintf(stderr, "\n");
    if (d_dbg.p) {   // -DMEGA_INSTR: paths finished per 0.25 ms of the last launch (x 32)
      unsigned long long h[128];
      WPT_CUDA(cudaMemcpy(h, d_dbg.p, sizeof h, cudaMemcpyDeviceToHost));
      std::fprintf(stderr, "wpt paths per 0.25 ms:");
      for (int i = 0; i < 100; i++) std::fprintf(stderr, " %llu", h[i] * 32ull);
      std::fprintf(stderr, "\n");
      // per-warp exit timeline: when did the warps exit relative to the first "queue empty", and what did the last ones hold
      const size_t nw = (size_t)device_sm_count() * 16 * 4;
      std::vector<unsigned long long> w(3 * nw);
      WPT_CUDA(cudaMemcpy(w.data(), d_dbg.p + 128, w.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      struct E { unsigned long long end, empty; uint32_t path, slot; };
      std::vector<E> ev;
      unsigned long long first_empty = ~0ull;
      for (size_t i = 0; i < nw; i++) if (w[3 * i]) { ev.push_back(E{w[3 * i], w[3 * i + 1], (uint32_t)(w[3 * i + 2] >> 32), (uint32_t)w[3 * i + 2]}); if (w[3 * i + 1]) first_empty = std::min(first_empty, w[3 * i + 1]); }
      std::sort(ev.begin(), ev.end(), [](const E& a, const E& b) { return a.end < b.end; });
      if (!ev.empty() && first_empty != ~0ull) {
        std::fprintf(stderr, "wpt warp exits after the first queue-empty (us), percentiles 10/50/90/99/100:");
        for (double q : {0.10, 0.50, 0.90, 0.99, 1.0}) { const E& e = ev[std::min(ev.size() - 1, (size_t)(q * (ev.size() - 1) + 0.5))]; std::fprintf(stderr, " %.0f", ((double)e.end - (double)first_empty) * 1e-3); }
        uint32_t mp = 0, ms = 0; for (const E& e : ev) { mp = std::max(mp, e.path); ms = std::max(ms, e.slot); }
        std::fprintf(stderr, "\nwpt longest path %u rays, longest slot %u rays (whole launch); the last 8 warps to exit (us after queue-empty, own longest path, own longest slot):", mp, ms);
        for (size_t i = ev.size() >= 8 ? ev.size() - 8 : 0; i < ev.size(); i++) std::fprintf(stderr, " (%.0f, %u, %u)", ((double)ev[i].end - (double)first_empty) * 1e-3, ev[i].path, ev[i].slot);
        std::fprintf(stderr, "\n");
      }
    }
  }
  out[0] = h_counters[0] + photon_rays; out[1] = h_counters[2]; out[2] = h_counters[1] + photon_visits;
  out[3] = photons_shot_total; out[4] = photons_stored_total; out[5] = iterations; out[6] = launches; out[7] = 0;
}

}  // namespace wpt
