// Session object behind the C ABI (include/wpt.h). One context = one GPU = one host thread
// at a time (the reference is single-threaded per instance, wasm_interface.rs:59-62).
#pragma once
#include <cuda_runtime.h>
#include <cstring>
#include <functional>
#include <map>
#include <list>
#include <string>
#include <vector>
#include "../../include/wpt.h"
#include "host_scene.h"
#include "kernels.h"

namespace wpt {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
void cuda_check(cudaError_t e, const char* what);
void nccl_unique_id(uint8_t out[128]);
#define WPT_CUDA(x) ::wpt::cuda_check((x), #x)

template <class T> struct DevBuf {
  T* p = nullptr; size_t n = 0;
  void alloc(size_t count) {
    if (count <= n && p) return;
    release();
    if (count) WPT_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
    n = count;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  void upload(const std::vector<T>& v, cudaStream_t s) {
    alloc(v.size());
    if (!v.empty()) WPT_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

struct HalfSettings { uint32_t type; uint32_t adaptive; };

struct Context {
  int device = 0;
  bool has_device = true;                  // false: host-only session (scene / BVH inspection)
  void require_device() const { if (!has_device) throw CudaError("no CUDA device in a host-only session (libwpt has no CPU fallback)"); }
  cudaStream_t stream = nullptr;       // stream all work is issued on
  cudaStream_t own_stream = nullptr;   // created by the session; `stream` may point elsewhere (set_stream)
  uint32_t W = 0, H = 0, scene_id = 0;
  float cam[5] = {0, 0, 0, 0, 0};
  wpt_config cfg;
  HalfSettings left{WPT_NORMAL_NEE, 0}, right{WPT_PNEE, 1};   // wasm_interface.rs:90-97
  uint32_t light_debug = 0;

  // meshes / textures (wasm_interface.rs:38-41)
  std::map<uint32_t, std::vector<float>> mesh_preload;
  std::map<uint32_t, std::vector<HostShape>> mesh_tris;
  std::map<uint32_t, std::vector<uint8_t>> textures;
  std::map<uint32_t, std::pair<uint32_t, uint32_t>> tex_dims;   // width, height per texture id
  DevBuf<uint8_t> d_tex[WPT_MAX_TEXTURES]; uint32_t d_tex_w[WPT_MAX_TEXTURES] = {0, 0, 0, 0}, d_tex_h[WPT_MAX_TEXTURES] = {0, 0, 0, 0};   // extension scene

  HostScene scene;
  DevBuf<DNode2> d_nodes2;
  DevBuf<DNode4> d_nodes4;
  DevBuf<DShape> d_shapes;
  DevBuf<DMaterial> d_mats;
  DevBuf<DLight> d_lights;

  // render targets (render_target.rs): accumulators + RGBA8 + sampling-debug RGBA8
  DevBuf<float4> d_accum;
  DevBuf<uint8_t> d_rgba, d_sampling;
  uint8_t* h_rgba = nullptr;       // pinned, W*H*4
  uint8_t* h_sampling = nullptr;   // pinned, W*H*4
  bool rgba_stale = true;

  // wavefront state
  DevBuf<float4> s_ray_o, s_ray_d, s_col, s_sh_o, s_sh_d, s_sh_c, s_tail;
  DevBuf<uint4> s_misc;
  DevBuf<float2> s_hit;
  DevBuf<uint32_t> s_pixel, s_spp;
  DevBuf<uint32_t> w_shadow_q0, w_shadow_q1, w_shadow_n, w_ring, w_work;
  DevBuf<unsigned long long> w_counters;
  uint32_t* h_ring = nullptr;                 // pinned [64]
  unsigned long long* h_counters = nullptr;   // pinned [8]
  uint32_t slots = 0;                         // slots of the current partition
  uint32_t slot_region[6] = {0, 0, 0, 0, 0, 0};   // region + rank/world the pixel map was built for
  DevBuf<uint32_t> s_pixel_alt, t_key, t_val, t_key2, t_val2; DevBuf<uint8_t> t_tmp;   // slot order by primary-hit class (launch_order_tiles)
  uint64_t slot_view = 0, scene_version = 0;   // camera / scene the slot order was computed for; bumped by upload_scene
  uint64_t view_key() const;

  // photons: records (loc.xyz, weight) + (light, shot), octree build scratch, flattened tree
  DevBuf<float4> ph_loc_w;
  DevBuf<uint2> ph_light_shot;
  DevBuf<uint32_t> ph_meta, ph_count, oc_node_of, oc_child_base, oc_count;
  DevBuf<unsigned long long> oc_fx;
  DevBuf<uint32_t> p_child_base, p_nbr;
  DevBuf<float> p_cum, p_bins;
  uint32_t p_nodes = 0;
  bool photons_ready = false;
  uint64_t photon_shots = 0, photon_count = 0;
  std::vector<uint32_t> ph_light; std::vector<float> ph_loc, ph_w;     // host copies (read-backs)
  std::vector<uint32_t> pt_meta; std::vector<float> pt_cum, pt_bins;   // flattened tree (read-backs)
  std::vector<uint32_t> oc_child_base_h, oc_count_h, oc_depth_h;       // host copy of the topology
  bool ph_host_valid = false, tree_host_valid = false;                 // read-back copies are made on first use
  void photon_list_host();
  void photon_tree_host();

  // pinned host copies of the flattened scene (re-upload without rebuilding)
  void* h_scene_blob = nullptr; size_t h_scene_bytes = 0;
  size_t blob_off[5] = {0, 0, 0, 0, 0}, blob_len[5] = {0, 0, 0, 0, 0};
  int64_t reupload_scene();

  // per-launch event timing (bench / profiling)
  bool profiling = false;
  struct EvPair { cudaEvent_t a, b; int kind; };
  std::vector<EvPair> ev_pending;
  std::vector<cudaEvent_t> ev_free;
  double prof_ms[6] = {0, 0, 0, 0, 0, 0}; uint64_t prof_n[6] = {0, 0, 0, 0, 0, 0};   // 0 path kernel, 1 shade kernel, 2 error map, 3 round render, 4 exchange, 5 photon warm-up
  cudaEvent_t ev_open[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void ev_mark(int kind, bool begin);   // profiling only: bracket a phase with events on the session's stream
  uint64_t adaptive_rounds = 0;
  double photon_build_ms = 0;             // wall clock of the last build_photons()
  unsigned long long prof_base[4] = {0, 0, 0, 0};
  unsigned long long life[4] = {0, 0, 0, 0};   // counters folded in by reset()
  void read_counters(unsigned long long out[4]);
  cudaEvent_t ev_get();
  void ev_harvest();
  void set_profiling(bool on);
  void profile_read(double out[8]);
  void profile_read_rounds(double out[8]);

  // sampling strategies (sampling_strategy.rs) per logical region
  struct Strategy {
    uint32_t rx = 0, ry = 0, rw = 0, rh = 0;
    DevBuf<uint32_t> round_left, take, round_spp;   // region-indexed
    DevBuf<float> mse;
    DevBuf<unsigned long long> state;   // device-driven adaptive rounds (kernels.cu k_ad_*)
    uint64_t left_total = 0;      // samples still queued in the current adaptive round
    bool started = false;         // the initial 4-spp queue has been issued
    bool painted = false;
    uint64_t random_ticks = 0;    // ticks consumed by the random strategy so far
    Strategy() = default;
    Strategy(const Strategy&) = delete;
  };
  std::list<Strategy> strategies;
  DevBuf<unsigned long long> a_stats, a_block_tot, a_block_suffix;
  std::function<void()> exchange_hook;   // multi-GPU: gather all rows' accumulators between adaptive rounds
  // native NCCL plane (dist_nccl.cpp): communicator + staging for the accumulator all-gather
  void* nccl_comm = nullptr; bool nccl_own = false;
  DevBuf<float4> x_send, x_recv;
  uint64_t collectives = 0;
  void attach_nccl(const uint8_t id128[128], void* existing_comm, uint32_t rank, uint32_t world);
  void detach_nccl();
  void exchange_native();
  void reduce_native(uint32_t* dev_words, uint64_t n);
  std::function<void(uint32_t*, uint64_t)> reduce_hook;   // multi-GPU: in-place sum-allreduce of 32-bit words on `stream` (photon batches)
  DevBuf<float4> d_seg_buf;    // k_wpool / k_mega: segment sums of a multi-segment launch (contract B10)
  DevBuf<unsigned long long> d_dbg;   // -DMEGA_INSTR diagnostics
  DevBuf<float4> d_pool;       // k_wpool: path contexts, 160 B each, pool_ctx per warp
  DevBuf<uint32_t> d_seg_cnt, d_seg_off, d_seg_list, d_pass_spp; DevBuf<uint8_t> d_scan_tmp;   // strategy rounds: segment list (contract B10)
  DevBuf<float4> d_seg_acc;    // wavefront engine: the running segment, per pixel
  DevBuf<uint32_t> ph_dense;   // one photon batch: [meta | light | loc_w x 4] per shot slot
  Strategy& strategy_for(uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh);
  void clear_strategies();
  void region_error_launch(Strategy& s);
  void region_error(Strategy& s, float stats3[3]);
  void render_take(Strategy& s, uint32_t render_type, bool bounded);
  uint64_t run_adaptive(Strategy& s, uint32_t render_type, uint64_t budget, const std::function<void()>& exchange);
  uint64_t run_adaptive_host(Strategy& s, uint32_t render_type, uint64_t budget, const std::function<void()>& exchange);
  void run_random(Strategy& s, uint32_t render_type, uint64_t ticks);
  void render_random(uint64_t ticks);

  // counters since the last reset
  uint64_t iterations = 0, launches = 0, photons_shot_total = 0, photons_stored_total = 0;
  uint64_t photon_rays = 0, photon_visits = 0;   // rays / node visits of the photon warm-up (counted on the host)

  Context(int device, uint32_t w, uint32_t h, uint32_t scene_id, const float cam5[5]);
  ~Context();

  void select_scene(uint32_t id);          // wasm_interface.rs:389-398 + upload
  void upload_scene();
  void alloc_targets();
  void clear_targets();
  void reset();
  void ensure_slots(uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh);
  void ensure_wavefront_state(uint32_t n);
  RenderParams params(uint32_t render_type) const;
  PathState path_state();
  WaveBuffers wave_buffers();
  void region(uint32_t* rx, uint32_t* ry, uint32_t* rw, uint32_t* rh) const;

  // the persistent kernels read at most two infinite planes from their parameters: scenes with more run on the wavefront engine
  bool use_wavefront() const { return cfg.engine == 1 || scene.num_inf > 2; }
  // drivers
  void run_wavefront(uint32_t render_type, const uint32_t* d_spp_per_slot, uint32_t uniform_spp);
  void run_persistent(uint32_t render_type, const uint32_t* d_spp_per_slot, uint32_t uniform_spp);
  void run_paths(uint32_t render_type, const uint32_t* d_spp_per_slot, uint32_t uniform_spp);   // engine switch
  void render_exact(uint32_t spp);
  uint64_t render_adaptive(uint64_t budget);
  void build_photons();
  void compute(uint64_t num_samples);      // wasm_interface.rs:374-384
  void photon_sample_batch(const float* pts3, const uint32_t* seeds, uint64_t n, uint32_t* light, float* pdf);
  void error_map(float* mse, float stats3[3]);
  void round_spp(uint32_t* spp);
  const uint8_t* results(uint32_t show_sampling);
  void stats(uint64_t out[8]);
};

}  // namespace wpt
