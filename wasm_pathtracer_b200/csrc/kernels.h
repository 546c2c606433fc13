// Launch wrappers for the sm_100a kernels (kernels.cu). Plain structs and pointers only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "wpt_types.h"

namespace wpt {

// slot flags (PathState::misc.z)
enum : uint32_t { SL_ACTIVE = 1, SL_BOUNCED = 2, SL_SHADOW = 4, SL_TAIL = 8, SL_DONE = 16, SL_FRESH = 32 };

// Wavefront state, SoA over slots. One slot owns one pixel and runs that pixel's samples
// one after the other (path regeneration), so per-pixel accumulation order is the sample
// order — bit-exact against the oracle and independent of scheduling.
struct PathState {
  float4* ray_o;      // origin.xyz, throughput.x
  float4* ray_d;      // dir.xyz,    throughput.y
  float4* col;        // colour.xyz, throughput.z
  uint4* misc;        // rng state, current sample index, flags, end sample index
  float2* hit;        // t, bits(shape id)  (id < 0: miss)
  float4* sh_o;       // shadow-ray origin.xyz, distance to the light point
  float4* sh_d;       // shadow-ray dir.xyz, bits(light shape id)
  float4* sh_c;       // contribution if unoccluded .xyz, bits(occluded) written by the trace kernel
  float4* tail;       // colour of a finished path that still waits for its last shadow ray
  uint32_t* pixel;    // viewport pixel index (y * W + x) of the slot
  uint32_t n;         // number of slots
};

struct RenderParams {
  DScene scene;
  DCamera cam;
  DPhotonTree photons;
  uint32_t W, H;
  uint32_t render_type;    // 0 NoNEE, 1 NormalNEE, 2 PNEE
  uint32_t light_debug;
  uint32_t base_seed;
};

// counters[0] = rays, [1] = node visits, [2] = paths finished, [3] = leaf primitive tests (u64 each)
struct WaveBuffers {
  uint32_t* shadow_q[2];      // slot indices with a shadow ray this iteration (double buffered)
  uint32_t* shadow_n;         // [2] queue lengths
  uint32_t* active_ring;      // [64] live-slot count per iteration (ring)
  unsigned long long* counters;
  float4* accum;              // per pixel: rgb sums, bits(sample count)
};

// Persistent path kernel (k_mega): slots are fetched from `work_counter`.
struct MegaParams {
  RenderParams rp;
  float4* accum;
  const uint32_t* pixel;          // slot -> viewport pixel index
  const uint32_t* spp_per_slot;   // may be null: `uniform_spp` for every slot
  uint32_t uniform_spp, nslots;       // nslots = pixel slots x nseg
  uint32_t nseg, seg_len;             // contract B10: a pixel's samples of this launch are cut into nseg segments of seg_len
  float4* seg_buf;                    // nseg > 1: segment sums [pixel slot * nseg + j], combined in order by launch_combine_segments
  uint32_t seg_buf_n;                 // entries of seg_buf (checked build: WPT_CHECK on every store)
  const uint32_t* seg_list;           // strategy rounds: slot -> (pixel slot << 3 | segment); then nslots comes from nslots_dev and seg_buf is indexed by slot
  const uint32_t* nslots_dev;
  uint32_t list_len;                  // strategy rounds cut into short slots of list_len samples (list entries pixel slot << 6 | slot, one colour per sample in seg_buf[slot * list_len ..]); 0 = segments of seg_len
  // render_exact, end zones of the slot queue (up to three, in queue order; unused ones start at 0xFFFFFFFF): zone z begins at slot zone_start[z]
  // with pixel slot zone_pslot[z]; a pixel has zone_per[z] slots of zone_len[z] samples. Zone slots write one colour per sample to
  // seg_buf[zone_samples + (pixel slot - zone_pslot[0]) * uniform_spp + sample].
  uint32_t zone_start[3], zone_pslot[3], zone_per[3], zone_len[3], zone_samples;
  uint32_t t_hi, t_lo;            // warp-vote thresholds of the traversal bursts
  uint32_t t_inner;               // leave the inner phase when lanes-at-inner * t_inner <= burst lanes
  uint32_t inner_reps;            // k_mega: inner-node steps per vote inside a burst
  uint32_t t_torus;               // k_mega: the deferred torus phase runs when this many lanes of the warp are parked at a torus (or nothing else is left)
  uint32_t t_switch;              // k_wpool: a logic run that cannot refill leaves when fewer than this many lanes still hold a context
  uint32_t t_refill;              // k_wpool: idle lanes of a traversal burst refill from the queue when at least this many are idle
  float4* pool;                   // k_wpool: path contexts, [warp][pool_ctx][10 x float4]
  uint32_t pool_ctx;              // k_wpool: contexts per warp (<= 128)
  uint32_t chunk;                 // slots a warp fetches at a time (multiple of 32)
  uint32_t scene_kind;            // kernel variant by scene content (device_core.cuh): 0 triangles + planes, 1 + tori + boxes (the reference's primitives), 2 + the extension
  float4 inf_q1[2];               // (normal, normal.location) of the (at most two) infinite planes (shape record q1)
  float4 root_a, root_b;          // the BVH2 root node (box + left_first, count)
  uint32_t seed_path;             // mix32(STREAM_PATH ^ base_seed): the constant part of a path's stream seed
  uint32_t* work_counter;         // zeroed before the launch
  unsigned long long* counters;
  unsigned long long* dbg;        // -DMEGA_INSTR: [0..99] paths finished per 0.25 ms of the launch (one lane in 32 sampled)
};
void launch_mega(const MegaParams& P, const int blocks_per_sm[4], cudaStream_t s);   // per kernel variant, see kernels.cu
// Warp-pool path kernel (k_wpool, wpool.cu): a warp owns a pool of path contexts and alternates between logic runs and traversal bursts.
void launch_wpool(const MegaParams& P, int grid, int blocks_per_sm, cudaStream_t s);
uint32_t wpool_warps(int grid);
size_t wpool_ctx_bytes();

void launch_setup_slots(const PathState& st, const uint32_t* spp_per_slot, uint32_t uniform_spp, const float4* accum, cudaStream_t s);
void launch_trace(const RenderParams& rp, const PathState& st, const WaveBuffers& wb, uint32_t iter, int grid, cudaStream_t s);
void launch_shade(const RenderParams& rp, const PathState& st, const WaveBuffers& wb, uint32_t iter, int grid, cudaStream_t s);
void launch_combine_segments(float4* accum, const uint32_t* pixel, uint32_t npix, const float4* seg_buf, uint32_t nseg, uint32_t spp, uint32_t zone_pslot, uint32_t seg_len, cudaStream_t s);
// Strategy rounds (per-slot sample counts): cut every pixel's samples into segments of seg_len, listed pixel by pixel.
// seg_off needs npix + 1 entries (seg_off[npix] = number of segments); scan_tmp is scratch of at least seg_scan_bytes(npix) bytes.
size_t seg_scan_bytes(uint32_t npix);
void launch_build_segment_list(const uint32_t* slot_spp, uint32_t npix, uint32_t seg_len, uint32_t* seg_cnt, uint32_t* seg_off, void* scan_tmp, size_t scan_bytes, uint32_t* seg_list, uint32_t shift, cudaStream_t s);
void launch_combine_segment_list(float4* accum, const uint32_t* pixel, uint32_t npix, const float4* seg_buf, const uint32_t* seg_off, const uint32_t* slot_spp, uint32_t list_len, uint32_t seg_len, cudaStream_t s);
void launch_segment_pass_spp(const uint32_t* slot_spp, uint32_t npix, uint32_t seg_len, uint32_t pass, uint32_t* pass_spp, uint32_t* any_left, cudaStream_t s);
void launch_add_segment(float4* accum, const uint32_t* pixel, uint32_t npix, const float4* seg_acc, cudaStream_t s);
void launch_clear_pixels(float4* buf, const uint32_t* pixel, uint32_t npix, cudaStream_t s);
void launch_resolve_rgba(const float4* accum, uint8_t* rgba, uint32_t n, cudaStream_t s);
void launch_primary_probe(const RenderParams& rp, int32_t* ids, uint32_t* visits, float* dist, cudaStream_t s);
void launch_trace_batch(const RenderParams& rp, const float* o, const float* d, uint64_t n, int32_t* ids, float* dist, uint32_t* visits, float* normals, cudaStream_t s);
// multi-GPU accumulator exchange: this rank's rows of the region -> contiguous staging (padded to `per` rows), and back from all ranks' staging
void launch_pack_rows(const float4* accum, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, uint32_t rank, uint32_t world, uint32_t per, float4* send, cudaStream_t s);
void launch_unpack_rows(const float4* recv, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, uint32_t rank, uint32_t world, uint32_t per, float4* accum, cudaStream_t s);
// slot order by primary-hit class (kernels.cu): tiles whose primary rays hit finite geometry first, background tiles last
size_t tile_sort_bytes(uint32_t ntiles);
void launch_order_tiles(const RenderParams& rp, const uint32_t* pixel, uint32_t n, uint32_t ntiles, uint32_t* key, uint32_t* val, uint32_t* key_out, uint32_t* val_out,
                        void* tmp, size_t tmp_bytes, uint32_t* pixel_out, cudaStream_t s);
void launch_fill_pixels(uint32_t* pixel, uint32_t W, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, uint32_t rank, uint32_t world, cudaStream_t s);

// photons
void launch_photon_emit(const RenderParams& rp, unsigned long long shot0, uint32_t n, uint32_t rank, uint32_t world, uint32_t* meta, uint32_t* rec_light, float4* rec_loc_w, cudaStream_t s);
void launch_photon_compact(const uint32_t* meta, const uint32_t* light, const float4* lw, uint32_t n, uint32_t remaining, uint32_t base, uint32_t rank, uint32_t world,
                           uint32_t* flag, uint32_t* off, void* scan_tmp, size_t scan_bytes, float4* out_lw, uint2* out_ls, unsigned long long* res, cudaStream_t s);
void launch_octree_assign(const float4* loc_w, uint32_t n, uint32_t* node_of, const uint32_t* child_base, uint32_t* count, cudaStream_t s);
void launch_octree_bins(const float4* loc_w, const uint2* light_shot, uint32_t n, const uint32_t* child_base, unsigned long long* fx, uint32_t num_lights, cudaStream_t s);
void launch_octree_cdf(const unsigned long long* fx, float* bins, float* cum, uint32_t num_nodes, uint32_t num_lights, cudaStream_t s);
void launch_photon_sample_batch(const DPhotonTree& t, const float* pts, const uint32_t* seeds, uint64_t n, uint32_t* light, float* pdf, cudaStream_t s);
// adaptive / random strategies
enum : unsigned long long { AD_IDLE = 0, AD_FIRST = 1, AD_ERR = 2, AD_CONT = 3 };   // mode of a device-driven adaptive step
void launch_error_map(const float4* accum, uint32_t W, uint32_t H, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, float* mse, unsigned long long* stats, const unsigned long long* gate, cudaStream_t s);
void launch_ad_setup(unsigned long long* st, unsigned long long budget, cudaStream_t s);
void launch_ad_begin(unsigned long long* st, cudaStream_t s);
void launch_ad_first(unsigned long long* st, uint32_t* round_left, uint32_t* round_spp, uint32_t n, cudaStream_t s);
void launch_ad_total(unsigned long long* st, cudaStream_t s);
void launch_ad_end(unsigned long long* st, cudaStream_t s);
void launch_cut_device(const uint32_t* left, uint32_t n, unsigned long long* block_tot, unsigned long long* block_suffix, const unsigned long long* room_dev, uint32_t* take, cudaStream_t s);
void launch_adaptive_spp(const float* mse, uint32_t n, unsigned long long* stats, uint32_t* round_left, uint32_t* round_spp, uint8_t* sampling_rgba8, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, const unsigned long long* gate, cudaStream_t s);
void launch_fill_region_rgba(uint8_t* rgba, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, uint32_t value, cudaStream_t s);
void launch_fill_u32(uint32_t* a, uint32_t n, uint32_t v, cudaStream_t s);
void launch_sum_u32(const uint32_t* a, uint32_t n, unsigned long long* out, cudaStream_t s);
void launch_cut(const uint32_t* left, uint32_t n, unsigned long long* block_tot, const unsigned long long* block_suffix, unsigned long long budget, uint32_t* take, int pass, cudaStream_t s);
void launch_gather_slot_spp(const uint32_t* take, const uint32_t* pixel, uint32_t nslots, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t* slot_spp, cudaStream_t s);
void launch_sub_u32(uint32_t* a, const uint32_t* b, uint32_t n, cudaStream_t s);
void launch_random_ticks(unsigned long long t0, unsigned long long n, uint32_t seed, uint32_t rw, uint32_t rh, uint32_t* take, cudaStream_t s);

int device_sm_count();

}  // namespace wpt
