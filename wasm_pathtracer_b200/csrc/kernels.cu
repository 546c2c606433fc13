// sm_100a kernels of the wavefront path tracer. See DESIGN.md for the stage diagram.
//
//   k_shade  : per slot — resolve last iteration's shadow ray, shade the extension hit
//              (tracer.rs:237-329), emit the bounce ray + optional shadow ray, Russian
//              roulette, accumulate finished paths and regenerate the next sample's
//              camera ray (tracer.rs:156-201).
//   k_trace  : persistent grid-stride kernel over [extension rays | shadow-ray queue]:
//              Scene::trace_g / Scene::shadow_ray (scene.rs:104-184).
#include "kernels.h"
#include "device_core.cuh"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

namespace wpt {

int device_sm_count() {   // of the current device (sessions on different GPUs may differ)
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = cache[dev & 63];
  if (!c) {
    cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
    if (c <= 0) c = 148;
  }
  return c;
}

#define TRACE_THREADS 128
#define SHADE_THREADS 128

}  // namespace wpt
#include "device_shade.cuh"
namespace wpt {

// ------------------------------------------------------------------ slot setup
__global__ void k_setup_slots(PathState st, const uint32_t* __restrict__ spp_per_slot, uint32_t uniform_spp, const float4* __restrict__ accum) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.n) return;
  uint32_t spp = spp_per_slot ? spp_per_slot[i] : uniform_spp;
  uint32_t s0 = __float_as_uint(accum[st.pixel[i]].w);   // samples already accumulated = next sample index
  st.misc[i] = make_uint4(0u, s0, spp ? SL_FRESH : SL_DONE, s0 + spp);
}
void launch_setup_slots(const PathState& st, const uint32_t* spp_per_slot, uint32_t uniform_spp, const float4* accum, cudaStream_t s) {
  if (!st.n) return;
  k_setup_slots<<<(st.n + 255) / 256, 256, 0, s>>>(st, spp_per_slot, uniform_spp, accum);
}

__global__ void k_fill_pixels(uint32_t* pixel, uint32_t W, uint32_t x0, uint32_t y0, uint32_t w, uint32_t rows, uint32_t rank, uint32_t world) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * rows) return;
  // slot order = 8x4 tiles over the full-tile area (a warp's 32 consecutive slots are one tile:
  // coherent primary rays, neighbouring surface points afterwards), then the ragged right and
  // bottom strips in raster order
  const uint32_t W8 = w & ~7u, R4 = rows & ~3u, nA = W8 * R4, nB = (w - W8) * R4;
  uint32_t row, col;
  if (i < nA) {
    uint32_t t = i >> 5, j = i & 31u, tiles_x = W8 >> 3;
    col = (t % tiles_x) * 8 + (j & 7u); row = (t / tiles_x) * 4 + (j >> 3);
  } else if (i < nA + nB) {
    uint32_t k = i - nA, sw = w - W8;
    row = k / sw; col = W8 + k % sw;
  } else {
    uint32_t k = i - nA - nB;
    row = R4 + k / w; col = k % w;
  }
  pixel[i] = (y0 + band_row(row, rank, world)) * W + x0 + col;
}
void launch_fill_pixels(uint32_t* pixel, uint32_t W, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, uint32_t rank, uint32_t world, cudaStream_t s) {
  uint32_t rows = band_rows(h, rank, world);
  if (!rows || !w) return;
  k_fill_pixels<<<(w * rows + 255) / 256, 256, 0, s>>>(pixel, W, x0, y0, w, rows, rank, world);
}

// ---- slot order by primary-hit class. A launch ends when its last path ends, and a lone lane runs a ray in ~25 us (a
// dependent chain of node fetches), so a 100-ray path that starts shortly before the slot queue runs dry keeps the whole GPU
// waiting for 2.5 ms (scripts/tail_probe.py). Long paths start on finite geometry; paths through background pixels are one ray
// long. So the tiles are dealt in three stably partitioned classes — (0) a primary ray hits a BVH shape, (1) only an infinite
// plane, (2) background — and the queue ends with the cheapest tiles: the long paths are over before the queue is empty. The
// order of the slots changes no result (per-path streams, per-slot segment sums, combined in order afterwards).
__global__ void k_tile_class(RenderParams rp, const uint32_t* __restrict__ pixel, uint32_t ntiles, uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, tile = gid >> 3, k = gid & 7u;
  const bool active = tile < ntiles;
  uint32_t cls = 2u;
  if (active) {   // eight probes per 8x4 tile, through the pixel centres
    const uint32_t pix = pixel[(size_t)tile * 32u + k * 4u + (k & 3u)];
    const uint32_t py = pix / rp.W, px = pix - py * rp.W;
    const Ray ray = camera_ray(rp.cam, px, py, 0.5f, 0.5f);
    const GHit g = trace_g(rp.scene, ray);
    cls = g.id < 0 ? 2u : ((uint32_t)g.id < rp.scene.num_inf ? 1u : 0u);
  }
  cls = min(cls, __shfl_xor_sync(0xFFFFFFFFu, cls, 1));
  cls = min(cls, __shfl_xor_sync(0xFFFFFFFFu, cls, 2));
  cls = min(cls, __shfl_xor_sync(0xFFFFFFFFu, cls, 4));
  if (active && k == 0u) { key[tile] = cls; val[tile] = tile; }
}
__global__ void k_permute_tiles(const uint32_t* __restrict__ in, const uint32_t* __restrict__ order, uint32_t ntiles, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (i >> 5) < ntiles ? in[(size_t)order[i >> 5] * 32u + (i & 31u)] : in[i];   // the ragged strips behind the full tiles stay where they are
}
size_t tile_sort_bytes(uint32_t ntiles) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)ntiles, 0, 2);
  return bytes;
}
void launch_order_tiles(const RenderParams& rp, const uint32_t* pixel, uint32_t n, uint32_t ntiles, uint32_t* key, uint32_t* val, uint32_t* key_out, uint32_t* val_out,
                        void* tmp, size_t tmp_bytes, uint32_t* pixel_out, cudaStream_t s) {
  if (!ntiles) return;
  k_tile_class<<<(ntiles * 8u + 127u) / 128u, 128, 0, s>>>(rp, pixel, ntiles, key, val);
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key_out, val, val_out, (int)ntiles, 0, 2, s);   // stable: raster order inside a class
  k_permute_tiles<<<(n + 255u) / 256u, 256, 0, s>>>(pixel, val_out, ntiles, n, pixel_out);
}

// ------------------------------------------------------------------ multi-GPU row exchange (band partition, wpt_types.h)
__global__ void k_pack_rows(const float4* __restrict__ accum, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rows, uint32_t rank, uint32_t world, float4* __restrict__ send) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * rw) return;
  uint32_t k = i / rw, col = i - k * rw;
  send[i] = accum[(size_t)(ry + band_row(k, rank, world)) * W + rx + col];
}
__global__ void k_unpack_rows(const float4* __restrict__ recv, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, uint32_t self, uint32_t world, uint32_t per, float4* __restrict__ accum) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)world * per * rw) return;
  uint32_t r = (uint32_t)(i / ((size_t)per * rw));
  uint32_t rem = (uint32_t)(i - (size_t)r * per * rw), k = rem / rw, col = rem - k * rw;
  if (r == self || k >= band_rows(rh, r, world)) return;   // own rows are in place; the padding of ranks with fewer rows
  accum[(size_t)(ry + band_row(k, r, world)) * W + rx + col] = recv[i];
}
void launch_pack_rows(const float4* accum, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, uint32_t rank, uint32_t world, uint32_t per, float4* send, cudaStream_t s) {
  (void)per;
  uint32_t rows = band_rows(rh, rank, world);
  if (!rows || !rw) return;
  k_pack_rows<<<(rows * rw + 255) / 256, 256, 0, s>>>(accum, W, rx, ry, rw, rows, rank, world, send);
}
void launch_unpack_rows(const float4* recv, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, uint32_t rank, uint32_t world, uint32_t per, float4* accum, cudaStream_t s) {
  size_t n = (size_t)world * per * rw;
  if (!n) return;
  k_unpack_rows<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(recv, W, rx, ry, rw, rh, rank, world, per, accum);
}

// ------------------------------------------------------------------ trace
__global__ void __launch_bounds__(TRACE_THREADS) k_trace(RenderParams rp, PathState st, WaveBuffers wb, uint32_t iter) {
  const uint32_t qsel = (iter + 1) & 1u;          // queue filled by the previous shade pass
  const uint32_t nshadow = wb.shadow_n[qsel];
  const uint32_t total = st.n + nshadow;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    wb.shadow_n[iter & 1u] = 0;                   // queue the next shade pass fills
    wb.active_ring[iter & 63u] = 0;
  }
  unsigned long long rays = 0, visits = 0, prims = 0;
  const uint32_t* __restrict__ sq = wb.shadow_q[qsel];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < st.n) {
      uint32_t flags = st.misc[i].z;
      if (!(flags & SL_ACTIVE)) continue;
      float4 o = st.ray_o[i], d = st.ray_d[i];
      Ray ray = make_ray(xyz(o), xyz(d));
      GHit g = trace_g(rp.scene, ray);
      st.hit[i] = make_float2(g.t, __int_as_float(g.id));
      rays += 1; visits += g.visits; prims += g.prims;
    } else {
      uint32_t slot = sq[i - st.n];
      float4 o = st.sh_o[slot], d = st.sh_d[slot];
      Ray ray = make_ray(xyz(o), xyz(d));
      GHit g = trace_g(rp.scene, ray);
      // Scene::shadow_ray, scene.rs:114-132
      bool occluded = g.id >= 0 && g.t < o.w && g.id != __float_as_int(d.w);
      st.sh_c[slot].w = __uint_as_float(occluded ? 1u : 0u);
      rays += 1; visits += g.visits; prims += g.prims;
    }
  }
  rays = warp_sum_u64(rays);
  visits = warp_sum_u64(visits);
  prims = warp_sum_u64(prims);
  __shared__ unsigned long long s_r[TRACE_THREADS / 32], s_v[TRACE_THREADS / 32], s_p[TRACE_THREADS / 32];
  if ((threadIdx.x & 31) == 0) { s_r[threadIdx.x >> 5] = rays; s_v[threadIdx.x >> 5] = visits; s_p[threadIdx.x >> 5] = prims; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long r = 0, v = 0, pc = 0;
    for (int k = 0; k < TRACE_THREADS / 32; k++) { r += s_r[k]; v += s_v[k]; pc += s_p[k]; }
    if (r) { atomicAdd(&wb.counters[0], r); atomicAdd(&wb.counters[1], v); atomicAdd(&wb.counters[3], pc); }
  }
}
void launch_trace(const RenderParams& rp, const PathState& st, const WaveBuffers& wb, uint32_t iter, int grid, cudaStream_t s) {
  k_trace<<<grid, TRACE_THREADS, 0, s>>>(rp, st, wb, iter);
}

// ------------------------------------------------------------------ shade
struct Accum {
  float4* acc;
  WPT_DEV void add(uint32_t pix, F3 c) {   // RenderTarget::write, render_target.rs:55-58
    float4 a = acc[pix];
    a.x += c.x; a.y += c.y; a.z += c.z;
    a.w = __uint_as_float(__float_as_uint(a.w) + 1u);
    acc[pix] = a;
  }
};

__global__ void __launch_bounds__(SHADE_THREADS) k_shade(RenderParams rp, PathState st, WaveBuffers wb, uint32_t iter) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = false;
  unsigned long long paths_done = 0;
  if (i < st.n) {
    uint4 m = st.misc[i];
    uint32_t flags = m.z;
    if (!(flags & SL_DONE)) {
      const uint32_t pix = st.pixel[i];
      Accum acc{wb.accum};
      float4 ro = st.ray_o[i], rd = st.ray_d[i], cl = st.col[i];
      F3 color = xyz(cl);
      F3 T = f3(ro.w, rd.w, cl.w);
      Rng rng; rng.s = m.x;
      // 1. last iteration's shadow ray (tracer.rs:299-308)
      if (flags & (SL_SHADOW | SL_TAIL)) {
        float4 c = st.sh_c[i];
        bool occluded = __float_as_uint(c.w) != 0u;
        if (flags & SL_TAIL) {
          F3 tc = xyz(st.tail[i]);
          if (!occluded) tc = tc + xyz(c);
          acc.add(pix, tc);
          paths_done++;
        } else if (!occluded) color = color + xyz(c);
        flags &= ~(SL_SHADOW | SL_TAIL);
      }
      bool need_regen = false, finished = false;
      if (flags & SL_FRESH) { need_regen = true; flags &= ~SL_FRESH; }
      else if (!(flags & SL_ACTIVE)) flags |= SL_DONE;   // drained: only a tail was pending
      else {
        // 2. shade the extension hit
        float2 h = st.hit[i];
        Ray ray = make_ray(xyz(ro), xyz(rd));
        PathRegs ps; ps.color = color; ps.T = T; ps.rng = rng; ps.bounced = (flags & SL_BOUNCED) != 0;
        ShadeOut so;
        shade_hit<K_EXT>(rp, ray, __float_as_int(h.y), h.x, ps, so);
        color = ps.color; T = ps.T; rng = ps.rng;
        if (so.finished) finished = true;
        else {
          ro.x = so.next_o.x; ro.y = so.next_o.y; ro.z = so.next_o.z;
          rd.x = so.next_d.x; rd.y = so.next_d.y; rd.z = so.next_d.z;
          if (ps.bounced) flags |= SL_BOUNCED; else flags &= ~SL_BOUNCED;   // a specular bounce (extension materials) clears it
          if (so.shadow) {
            st.sh_o[i] = make_float4(so.sh_o.x, so.sh_o.y, so.sh_o.z, so.sh_len);
            st.sh_d[i] = make_float4(so.sh_d.x, so.sh_d.y, so.sh_d.z, __int_as_float(so.sh_light));
            st.sh_c[i] = make_float4(so.contrib.x, so.contrib.y, so.contrib.z, 0.0f);
            uint32_t q = atomicAdd(&wb.shadow_n[iter & 1u], 1u);
            wb.shadow_q[iter & 1u][q] = i;
            if (so.survive) flags |= SL_SHADOW;
            else { st.tail[i] = make_float4(color.x, color.y, color.z, 0.0f); flags |= SL_TAIL; need_regen = true; }
          } else if (!so.survive) finished = true;
        }
      }
      if (finished) { acc.add(pix, color); paths_done++; need_regen = true; }
      if (need_regen) {
        uint32_t s = (flags & SL_ACTIVE) ? m.y + 1 : m.y;   // FRESH slots start at m.y
        flags &= ~(SL_ACTIVE | SL_BOUNCED);
        if (s < m.w) {
          // tracer.rs:176-196 — sample s of this pixel on its own stream
          rng.s = stream_seed(pix, s, STREAM_PATH, rp.base_seed);
          float j1 = rng.f32();
          float j2 = rng.f32();
          uint32_t py = pix / rp.W, px = pix - py * rp.W;
          Ray cr = camera_ray(rp.cam, px, py, j1, j2);
          ro = make_float4(cr.o.x, cr.o.y, cr.o.z, 1.0f);
          rd = make_float4(cr.d.x, cr.d.y, cr.d.z, 1.0f);
          color = f3(0, 0, 0); T = f3(1.0f, 1.0f, 1.0f);
          flags |= SL_ACTIVE;
        } else if (!(flags & SL_TAIL)) flags |= SL_DONE;
        m.y = s;
      }
      ro.w = T.x; rd.w = T.y;
      st.ray_o[i] = ro; st.ray_d[i] = rd;
      st.col[i] = make_float4(color.x, color.y, color.z, T.z);
      m.x = rng.s; m.z = flags;
      st.misc[i] = m;
      live = !(flags & SL_DONE);
    }
  }
  // live-slot count for the host's termination test; finished-path counter
  unsigned int ballot = __ballot_sync(0xFFFFFFFFu, live);
  paths_done = warp_sum_u64(paths_done);
  if ((threadIdx.x & 31) == 0) {
    if (ballot) atomicAdd(&wb.active_ring[iter & 63u], (uint32_t)__popc(ballot));
    if (paths_done) atomicAdd(&wb.counters[2], paths_done);
  }
}
void launch_shade(const RenderParams& rp, const PathState& st, const WaveBuffers& wb, uint32_t iter, int grid, cudaStream_t s) {
  (void)grid;
  if (!st.n) return;
  k_shade<<<(st.n + SHADE_THREADS - 1) / SHADE_THREADS, SHADE_THREADS, 0, s>>>(rp, st, wb, iter);
}

// ------------------------------------------------------------------ persistent path kernel
// k_mega: the same path tracer as k_trace + k_shade, but the wavefront lives in registers.
// Every lane owns one pixel at a time and runs its samples one after the other (so the
// per-pixel accumulation order is the sample order, as in the wavefront engine); a lane is
// always in one of two stages — LOGIC (consume a finished trace: shade / resolve the shadow
// ray / finish the sample; generate the next camera ray; start the next trace with the plane
// tests and the root guard) or TRAV (one BVH node per step). The warp votes with __ballot_sync
// which stage to run: traversal bursts start when at least MEGA_T_HI lanes wait in TRAV (or no
// lane has logic to do) and stop when fewer than MEGA_T_LO are left, then the idle lanes refill
// themselves through LOGIC — the warp-ballot refill of Aila & Laine's persistent while-while
// kernel. Rays that end at the root guard (most rays of the bunny scene) never enter a burst.
#define MEGA_THREADS 128
#ifndef MEGA_T_HI
#define MEGA_T_HI 20
#endif
#ifndef MEGA_T_LO
#define MEGA_T_LO 10
#endif
enum : int { PH_NEED = 0, PH_LOGIC = 1, PH_TRAV = 2, PH_DONE = 3, PH_TORUS = 4 };
enum : int { ST_GEN = 0, ST_EXTEND = 1, ST_SHADOW = 2 };

// Scene::trace_g up to the BVH for the triangles / planes variant (scene.rs:162-212): at most two infinite shapes, all planes
// (plane.rs:80-99), and the root guard, all read from the kernel parameters (constant bank) instead of memory. Same tests in
// the same order as trav_begin: the first plane hit is accepted as is (scene.rs:438-440), a later one needs 0 < t < best.
template <int BVH>
WPT_DEV bool trav_begin_const(const MegaParams& P, const Ray& ray, Trav& tv) {
  bool have = false; float it = 0.0f; int iid = -1;
#pragma unroll
  for (int i = 0; i < 2; i++) {
    if ((uint32_t)i < P.rp.scene.num_inf) {
      const float4 q1 = P.inf_q1[i];
      const F3 nr = xyz(q1);
      const float n_dot_dir = dot(nr, ray.d);
      const float t = (q1.w - dot(nr, ray.o)) / n_dot_dir;
      const bool ok = n_dot_dir != 0.0f && t > 0.0f && (have ? t < it : t <= WPT_INF);
      if (ok) { have = true; it = t; iid = i; }
    }
  }
  tv.inf_t = it; tv.inf_id = iid;
  tv.bound = have ? it : WPT_INF;
  tv.best_id = -1;
  tv.visits = 0; tv.prims = 0; tv.sp = 0; tv.lcur = 0u;
  if (BVH == 4) { tv.lf = 0u; tv.cnt = 0u; return true; }
  const float4 ra = P.root_a, rb = P.root_b;
  tv.visits = 1;
  float h;
  if (!(box_hit(ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, ray, &h) && h < tv.bound)) return false;
  tv.lf = __float_as_uint(rb.z); tv.cnt = __float_as_uint(rb.w);
  return true;
}

template <int BVH, int KIND, int MINB, int RT, int ZN>
__global__ void __launch_bounds__(MEGA_THREADS, MINB) k_mega(MegaParams P) {
  const DScene& sc = P.rp.scene;
  uint32_t stack_n[WPT_STACK]; float stack_d[WPT_STACK];
  const unsigned FULL = 0xFFFFFFFFu;
#ifdef MEGA_NO_DEFER_TORUS
  constexpr bool DEFER_TORUS = false;
#else
  constexpr bool DEFER_TORUS = KIND != K_SIMPLE;   // variants whose scenes can hold tori
#endif
  // Ray::new's reciprocal direction (ray.rs:31-33): computed where the ray is made (three sites + the camera), or once per half
  // of a logic pass for every ray that starts there. One site runs with more lanes but keeps three more values live — measured
  // (gpurun_out/r2g_ab.log): museum 33.3 -> 31.8 ms, but headline 16.9 -> 17.2 and BVH4 + PNEE 30.8 -> 32.1 ms.
  constexpr bool INV_ONE_SITE = KIND != K_SIMPLE;
  // end zones of the slot queue (run_persistent): only in the variants that gain from them — their code costs the triangles /
  // planes variants 1.5 % (headline) to 6 % (BVH4 + PNEE) even when unused (gpurun_out/r2k_ab.log)
  // ZN = 0: segments only; 1: + end zones of render_exact (KIND != K_SIMPLE); 2: + strategy rounds in short slots (P.list_len != 0, any KIND).
  // Separate instantiations: code that a launch does not use still costs these kernels 1.5 - 6 % (register budget).
  // 3: a strategy round in segments (the list, no zones). 0 / 1 are render_exact launches (no list), 2 / 3 strategy rounds (list).
  constexpr bool ZONES = ZN == 1 || ZN == 2, SHORT_LIST = ZN == 2, LIST = ZN >= 2;
  const unsigned lane = threadIdx.x & 31u;
  int phase = PH_NEED, what = ST_GEN;
  uint32_t pixp = 0, s = 0, s_end = 0;   // pixp = px | py << 16 of the slot's pixel
  PathRegs ps; ps.color = f3(0, 0, 0); ps.T = f3(1, 1, 1); ps.rng.s = 1u; ps.bounced = false;
  Ray ray = make_ray(f3(0, 0, 0), f3(1, 1, 1));
  auto set_ray = [&](F3 o, F3 d) { if (INV_ONE_SITE) { ray.o = o; ray.d = d; } else ray = make_ray(o, d); };
  Trav tv; tv.lf = tv.cnt = 0; tv.sp = 0; tv.bound = 0; tv.best_id = -1; tv.inf_t = 0; tv.inf_id = -1; tv.visits = tv.prims = 0; tv.lcur = 0u;
  // The state a lane keeps across its shadow ray (bounce origin / direction, the light's contribution): nine floats that are written
  // once and read once per shadow ray. The BVH2 triangles + planes variants keep them in shared memory ([k][thread]: no bank conflicts) —
  // spills 242 -> 178 B per thread, headline 16.73 -> 16.59 ms, BVH2 + PNEE 24.2 -> 23.9 ms; BVH4 + PNEE unchanged and the museum variant
  // slower (29.1 -> 31.1 ms: its L1 shrinks by the 37 KB), so those keep registers (gpurun_out/r2e_smem.log).
  constexpr bool EXT_SMEM = KIND == K_SIMPLE && BVH == 2;
  __shared__ float sh_ext[EXT_SMEM ? 9 * MEGA_THREADS : 1];
  F3 ext_o = f3(0, 0, 0), ext_d = f3(0, 0, 0), contrib = f3(0, 0, 0);
  auto ext_put = [&](F3 o, F3 d, F3 c) {
    if (EXT_SMEM) {
      float* q = sh_ext + threadIdx.x;
      q[0] = o.x; q[MEGA_THREADS] = o.y; q[2 * MEGA_THREADS] = o.z; q[3 * MEGA_THREADS] = d.x; q[4 * MEGA_THREADS] = d.y; q[5 * MEGA_THREADS] = d.z;
      q[6 * MEGA_THREADS] = c.x; q[7 * MEGA_THREADS] = c.y; q[8 * MEGA_THREADS] = c.z;
    } else { ext_o = o; ext_d = d; contrib = c; }
  };
  auto ext_get = [&](int k) {
    if (EXT_SMEM) { const float* q = sh_ext + threadIdx.x + 3 * k * MEGA_THREADS; return f3(q[0], q[MEGA_THREADS], q[2 * MEGA_THREADS]); }
    return k == 0 ? ext_o : (k == 1 ? ext_d : contrib);
  };
  float sh_len = 0.0f; int sh_light = -1; bool alive_after_shadow = false;
  // ray / visit / primitive-test / path counters: warp sums (ballot + redux.sync) in uniform registers, one set of global
  // atomics per warp at the end — no per-ray atomics
  uint32_t n_rays = 0, n_visits = 0, n_paths = 0, n_prims = 0;
  uint32_t chunk_next = 0, chunk_end = 0, spare_next = 0, spare_end = 0; bool queue_empty = false;   // warp-uniform
  F3 acc_rgb = f3(0, 0, 0);   // sum of this segment's samples (contract B10), added to the pixel's accumulator (render_target.rs:8) when the segment is done
  uint32_t slot_id = 0;
#ifdef MEGA_INSTR
  unsigned long long i_lp = 0, i_ll = 0, i_ts = 0, i_tl = 0, i_sh = 0;   // logic passes, logic lanes, trav steps, trav lanes, shade lanes
  unsigned long long t_start, t_empty = 0;   // %globaltimer (ns): launch timeline = first start .. first "queue empty" .. last exit (scripts/tail_probe.py)
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
  uint32_t i_path_rays = 0, i_slot_rays = 0, i_max_path = 0, i_max_slot = 0;   // rays of the lane's current path / slot, and their maxima
#endif

  const uint32_t nslots = LIST ? *P.nslots_dev : P.nslots;
  for (;;) {
    // ---- pixel fetch: the warp owns a chunk of consecutive slots (= neighbouring tiles) and
    // hands them to its lanes; one atomic per chunk
    unsigned need = __ballot_sync(FULL, phase == PH_NEED);
    if (need) {
      uint32_t want = (uint32_t)__popc(need);
      if (chunk_next + want > chunk_end && !queue_empty) {
        // not enough left: the rest of the chunk is used first, then a new chunk is fetched
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(P.work_counter, P.chunk);
        base = __shfl_sync(FULL, base, 0);
        if (chunk_next >= chunk_end) { chunk_next = base; chunk_end = min(base + P.chunk, nslots); if (base >= nslots) { chunk_end = chunk_next = nslots; queue_empty = true; } }
        else { spare_next = base; spare_end = min(base + P.chunk, nslots); if (base >= nslots) { spare_next = spare_end = nslots; queue_empty = true; } }
#ifdef MEGA_INSTR
        if (queue_empty && !t_empty) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_empty));
#endif
      }
      if (phase == PH_NEED) {
        uint32_t r = (uint32_t)__popc(need & ((1u << lane) - 1u));
        uint32_t avail = chunk_end - chunk_next;
        uint32_t idx = r < avail ? chunk_next + r : spare_next + (r - avail);
        bool ok = r < avail ? true : (idx < spare_end);
        if (ok) {
          // contract B10: the samples of this launch are summed per segment from +0; a slot is one segment
          uint32_t pslot = idx, j = 0, zlen = 0;
          if (LIST) {   // strategy round: (pixel slot, segment) from the list; with P.list_len (pixel slot, short slot of list_len samples)
            const uint32_t e = P.seg_list[idx];
            if (SHORT_LIST) { pslot = e >> 6; j = e & 63u; zlen = P.list_len; } else { pslot = e >> 3; j = e & 7u; }
          }
          else if (ZN == 1 && idx >= P.zone_start[0]) {   // end zones of the queue: shorter slots (zone_len samples), see run_persistent
            const int z = idx >= P.zone_start[2] ? 2 : (idx >= P.zone_start[1] ? 1 : 0);
            const uint32_t r = idx - P.zone_start[z], q = r / P.zone_per[z];
            pslot = P.zone_pslot[z] + q; j = r - q * P.zone_per[z]; zlen = P.zone_len[z];
          }
          else if (P.nseg > 1) { pslot = idx / P.nseg; j = idx - pslot * P.nseg; }
          slot_id = idx;
          WPT_CHECK(idx < nslots);
          const uint32_t pix = P.pixel[pslot];
          WPT_CHECK(pix < P.rp.W * P.rp.H);
          const uint32_t py = pix / P.rp.W;   // once per slot, not per sample
          pixp = (pix - py * P.rp.W) | (py << 16);
          uint32_t spp = LIST ? P.spp_per_slot[pslot] : P.uniform_spp;
          uint32_t s0 = __float_as_uint(P.accum[pix].w);   // samples accumulated so far = next sample index
          const uint32_t len = zlen ? zlen : P.seg_len, b = min(j * len, spp), e = min(b + len, spp);
          s = s0 + b;
          s_end = s0 + e;
          // a zone slot stores every sample's colour on its own (k_combine_segments forms the segment sums): slot_id = flag | index one past its last sample
          if (ZONES && zlen) slot_id = 0x80000000u | (SHORT_LIST ? idx * zlen + (e - b) : P.zone_samples + (pslot - P.zone_pslot[0]) * P.uniform_spp + e);   // list: slot idx owns entries idx * len ..
          acc_rgb = f3(0.0f, 0.0f, 0.0f);
#ifdef MEGA_INSTR
          i_slot_rays = 0;
#endif
          what = ST_GEN; phase = PH_LOGIC;
        } else if (queue_empty) phase = PH_DONE;
      }
      {
        uint32_t avail = chunk_end - chunk_next;
        if (want <= avail) chunk_next += want;
        else {
          uint32_t from_spare = min(want - avail, spare_end - spare_next);
          chunk_next = spare_next + from_spare; chunk_end = spare_end; spare_next = spare_end = 0;
        }
      }
    }
    // a finished bounce / camera ray whose winner is a torus: its normal comes from the torus phase too (shade mode of trav_torus)
    if (DEFER_TORUS && phase == PH_LOGIC && what == ST_EXTEND && tv.best_id >= 0 && (tv.lcur & (LC_BEST_TORUS | LC_NREADY)) == LC_BEST_TORUS) { phase = PH_TORUS; tv.lcur |= LC_SHADE; }
    unsigned trav = __ballot_sync(FULL, phase == PH_TRAV);
    unsigned logic = __ballot_sync(FULL, phase == PH_LOGIC);
    if (DEFER_TORUS) {
      // ---- deferred torus phase (scenes with tori): a lane whose leaf scan meets a torus that passes the conservative cull
      // parks in PH_TORUS; the f64 quartic (~37 KB of code) runs for all parked lanes of the warp at once, when P.t_torus of
      // them wait or nothing else is left to do — many lanes per pass and few passes, instead of ~3 lanes whenever a leaf
      // step happens to meet a torus. Same arithmetic and the same ordered acceptance as leaf_scan.
      const unsigned tor = __ballot_sync(FULL, phase == PH_TORUS);
      if (!(trav | logic | tor)) break;
      if (tor && (__popc(tor) >= (int)P.t_torus || !(trav | logic))) {
        if (phase == PH_TORUS) {
          phase = (tv.lcur & LC_SHADE) ? PH_LOGIC : PH_TRAV;
          if (trav_torus<BVH>(sc, ray, tv) && !trav_pop<BVH>(sc, tv, stack_n, stack_d)) phase = PH_LOGIC;
        }
        continue;
      }
    }
    if (!(trav | logic)) break;
    if (__popc(trav) >= (int)P.t_hi || !logic) {
      // ---- traversal burst, while-while: one step per iteration — an inner-node step while enough of the burst's lanes
      // are at inner nodes, else a leaf step with every lane that waits at a leaf (triangle tests are the expensive
      // body: run them with as many lanes as possible), then one shared pop
      do {
        const bool tr = phase == PH_TRAV;
        const bool leaf = tr && trav_at_leaf<BVH>(tv);
        const int n_inner = __popc(__ballot_sync(FULL, tr && !leaf));
        bool need_pop = false;
        if (n_inner == __popc(trav) || n_inner * (int)P.t_inner > __popc(trav)) {
#ifdef MEGA_INSTR
          i_ts += 1; i_tl += n_inner;
#endif
          // a vote costs about as much as a third of an inner step: descend P.inner_reps nodes per vote (lanes that
          // reach a leaf or finish sit the rest out)
#pragma unroll 1
          for (uint32_t rep = 0; rep < P.inner_reps; rep++) {
            bool np = false;
            if (phase == PH_TRAV && !trav_at_leaf<BVH>(tv)) np = trav_inner<BVH>(sc, ray, tv, stack_n, stack_d);
            if (np && !trav_pop<BVH>(sc, tv, stack_n, stack_d)) phase = PH_LOGIC;
          }
        } else {
          if (leaf) {
            if (DEFER_TORUS) { if (trav_leaf_deferred<BVH, KIND>(sc, ray, tv)) phase = PH_TORUS; else need_pop = true; }
            else { trav_leaf<BVH, KIND>(sc, ray, tv); need_pop = true; }
          }
          if (need_pop && !trav_pop<BVH>(sc, tv, stack_n, stack_d)) phase = PH_LOGIC;
        }
        trav = __ballot_sync(FULL, phase == PH_TRAV);
      } while (__popc(trav) >= (int)P.t_lo);
      continue;
    }
#ifdef MEGA_INSTR
    i_lp += 1; i_ll += __popc(logic); i_sh += __popc(__ballot_sync(FULL, phase == PH_LOGIC && what == ST_EXTEND));
#endif
    // ---- logic pass, two halves. Half 0: lanes holding a shadow-ray result resolve it and start
    // their bounce ray; half 1: every lane holding a bounce/camera-ray result — including those
    // whose ray of half 0 ended at the root guard — is shaded. Both halves regenerate finished
    // lanes. So all lanes of the warp meet in the (expensive) shading code once per pass instead
    // of alternating shade / shadow-resolve in two populations that never line up.
    // (MEGA_ONE_BEGIN: one pass = every LOGIC lane consumes one result — shadow or hit — and starts at most one ray,
    //  so the plane tests + root guard have a single code site.)
#ifdef MEGA_ONE_BEGIN
    {
      const bool cons = phase == PH_LOGIC && what != ST_GEN;
#else
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      const bool cons = phase == PH_LOGIC && what == (half == 0 ? ST_SHADOW : ST_EXTEND);
#endif
      bool start = false, finish = false;
      GHit g; g.t = 0.0f; g.id = -1; g.visits = 0; g.prims = 0;
      if (cons) g = trav_result(tv);
      {
        const unsigned mc = __ballot_sync(FULL, cons);
        if (mc) { n_rays += (uint32_t)__popc(mc); n_visits += __reduce_add_sync(FULL, g.visits); n_prims += __reduce_add_sync(FULL, g.prims); }
      }
#ifdef MEGA_INSTR
      if (cons) { i_path_rays++; i_slot_rays++; i_max_path = max(i_max_path, i_path_rays); i_max_slot = max(i_max_slot, i_slot_rays); }
#endif
      if (cons) {
        if (what == ST_SHADOW) {   // Scene::shadow_ray, scene.rs:114-132
          bool occluded = g.id >= 0 && g.t < sh_len && g.id != sh_light;
          if (!occluded) ps.color = ps.color + ext_get(2);
          if (alive_after_shadow) { set_ray(ext_get(0), ext_get(1)); what = ST_EXTEND; start = true; }
          else finish = true;
        } else {
          ShadeOut so;
          if (DEFER_TORUS) { const TorusPre pre = torus_pre(tv); shade_hit<KIND, RT, true>(P.rp, ray, g.id, g.t, ps, so, &pre); }
          else shade_hit<KIND, RT>(P.rp, ray, g.id, g.t, ps, so);
          if (so.finished) finish = true;
          else if (so.shadow) {
            ext_put(so.next_o, so.next_d, so.contrib); sh_len = so.sh_len; sh_light = so.sh_light;
            alive_after_shadow = so.survive;
            set_ray(so.sh_o, so.sh_d); what = ST_SHADOW; start = true;
          } else if (so.survive) { set_ray(so.next_o, so.next_d); what = ST_EXTEND; start = true; }
          else finish = true;
        }
        if (finish) {   // RenderTarget::write, render_target.rs:55-58
          if (ZONES && (slot_id >> 31)) { WPT_CHECK((slot_id & 0x7FFFFFFFu) - (s_end - s) < P.seg_buf_n); P.seg_buf[(slot_id & 0x7FFFFFFFu) - (s_end - s)] = make_float4(ps.color.x, ps.color.y, ps.color.z, 0.0f); }
          else acc_rgb = acc_rgb + ps.color;
          s += 1; what = ST_GEN;
        }
#ifdef MEGA_INSTR
        if (finish && lane == 0 && P.dbg) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); unsigned long long bkt = (t - t_start) / 250000ull; atomicAdd(&P.dbg[bkt < 99 ? bkt : 99], 1ull); }
#endif
      }
      n_paths += (uint32_t)__popc(__ballot_sync(FULL, finish));
      if (phase == PH_LOGIC && what == ST_GEN) {
        const uint32_t px = pixp & 0xFFFFu, py = pixp >> 16, pix = py * P.rp.W + px;
        if (s < s_end) {   // tracer.rs:176-196 — sample s of this pixel on its own stream (stream_seed with the constant part from the host)
          uint32_t sd = mix32(pix + mix32(s + P.seed_path));
          ps.rng.s = sd == 0 ? 0xBABABEBEu : sd;
          float j1 = ps.rng.f32();
          float j2 = ps.rng.f32();
          set_ray(f3(P.rp.cam.ox, P.rp.cam.oy, P.rp.cam.oz), camera_dir(P.rp.cam, px, py, j1, j2));
          ps.color = f3(0, 0, 0); ps.T = f3(1.0f, 1.0f, 1.0f); ps.bounced = false;
#ifdef MEGA_INSTR
          i_path_rays = 0;
#endif
          what = ST_EXTEND; start = true;
        } else {
          if (ZONES && (slot_id >> 31)) { }
          else if (ZONES ? P.seg_buf != nullptr : (LIST || P.nseg > 1)) { WPT_CHECK(slot_id < P.seg_buf_n); P.seg_buf[slot_id] = make_float4(acc_rgb.x, acc_rgb.y, acc_rgb.z, 0.0f); }
          else {   // the only segment of its pixel: add it here
            float4 a = P.accum[pix];
            P.accum[pix] = make_float4(a.x + acc_rgb.x, a.y + acc_rgb.y, a.z + acc_rgb.z, __uint_as_float(s));
          }
          phase = PH_NEED;
        }
      }
      if (start) {
        if (INV_ONE_SITE) ray.inv = f3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
        if (trav_begin_const<BVH>(P, ray, tv)) phase = PH_TRAV;
      }
    }
  }
  // ---- counters
  if (lane == 0) {
    if (n_rays) atomicAdd(&P.counters[0], (unsigned long long)n_rays);
    if (n_visits) atomicAdd(&P.counters[1], (unsigned long long)n_visits);
    if (n_paths) atomicAdd(&P.counters[2], (unsigned long long)n_paths);
    if (n_prims) atomicAdd(&P.counters[3], (unsigned long long)n_prims);
  }
#ifdef MEGA_INSTR
  if (lane == 0) { atomicAdd(&P.counters[4], i_lp); atomicAdd(&P.counters[5], i_ll); atomicAdd(&P.counters[6], i_ts); atomicAdd(&P.counters[7], i_tl); }
  if (lane == 1) atomicAdd(&P.counters[8], i_sh);
  if (lane == 2) {
    unsigned long long t_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    atomicMin(&P.counters[9], t_start); if (t_empty) atomicMin(&P.counters[10], t_empty); atomicMax(&P.counters[11], t_end);
    atomicAdd(&P.counters[12], t_end - t_start);   // sum over warps of their lifetimes
  }
  {   // per warp: exit time, time it saw the queue empty, longest path and longest slot (rays) of its lanes
    const uint32_t mp = __reduce_max_sync(FULL, i_max_path), ms = __reduce_max_sync(FULL, i_max_slot);
    if (lane == 3 && P.dbg) {
      unsigned long long t_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
      const uint32_t w = blockIdx.x * (MEGA_THREADS / 32) + (threadIdx.x >> 5);
      P.dbg[128 + 3 * w] = t_end; P.dbg[128 + 3 * w + 1] = t_empty; P.dbg[128 + 3 * w + 2] = ((unsigned long long)mp << 32) | ms;
    }
  }
#endif
}
// Register budgets (blocks of 128 threads per SM) instantiated per variant: the measured optimum (gpurun_out/sweep8.log,
// sweep11.log, sweep12.log) — triangles/planes BVH2 8 (64 registers), BVH4 5 and 8, ext 16 (32 registers); the reference-primitives
// variant (tori, boxes) has 8 / 12 / 16 selectable at run time (WPT_MEGA_MINBG). With -DWPT_TUNING: 4 / 5 / 8 / 12 / 16 for every variant.
template <int BVH, int KIND, int RT, int ZN>
static void launch_mega_z(const MegaParams& P, int blocks_per_sm, cudaStream_t s) {
  auto grid_for = [&](int b) { int grid = device_sm_count() * b, need = (int)((P.nslots + MEGA_THREADS - 1) / MEGA_THREADS); return (grid > need && !P.nslots_dev) ? need : grid; };
#ifdef WPT_TUNING
  int b = blocks_per_sm >= 16 ? 16 : blocks_per_sm >= 12 ? 12 : blocks_per_sm >= 8 ? 8 : blocks_per_sm >= 5 ? 5 : 4;
  switch (b) {
    case 5: k_mega<BVH, KIND, 5, RT, ZN><<<grid_for(5), MEGA_THREADS, 0, s>>>(P); break;
    case 8: k_mega<BVH, KIND, 8, RT, ZN><<<grid_for(8), MEGA_THREADS, 0, s>>>(P); break;
    case 12: k_mega<BVH, KIND, 12, RT, ZN><<<grid_for(12), MEGA_THREADS, 0, s>>>(P); break;
    case 16: k_mega<BVH, KIND, 16, RT, ZN><<<grid_for(16), MEGA_THREADS, 0, s>>>(P); break;
    default: k_mega<BVH, KIND, 4, RT, ZN><<<grid_for(4), MEGA_THREADS, 0, s>>>(P); break;
  }
#else
  if constexpr (KIND == K_EXT) k_mega<BVH, KIND, 16, RT, ZN><<<grid_for(16), MEGA_THREADS, 0, s>>>(P);
  else if constexpr (KIND == K_REF) {
    if (BVH == 4 || blocks_per_sm >= 16) k_mega<BVH, KIND, 16, RT, ZN><<<grid_for(16), MEGA_THREADS, 0, s>>>(P);
    else if constexpr (BVH == 2) {
      if (blocks_per_sm >= 12) k_mega<BVH, KIND, 12, RT, ZN><<<grid_for(12), MEGA_THREADS, 0, s>>>(P);
      else k_mega<BVH, KIND, 8, RT, ZN><<<grid_for(8), MEGA_THREADS, 0, s>>>(P);
    }
  }
  else if constexpr (BVH == 2 || RT == 2) k_mega<BVH, KIND, 8, RT, ZN><<<grid_for(8), MEGA_THREADS, 0, s>>>(P);
  else k_mega<BVH, KIND, 5, RT, ZN><<<grid_for(5), MEGA_THREADS, 0, s>>>(P);
#endif
}
template <int BVH, int KIND, int RT>
static void launch_mega_t(const MegaParams& P, int blocks_per_sm, cudaStream_t s) {
  if (P.seg_list) { if (P.list_len) launch_mega_z<BVH, KIND, RT, 2>(P, blocks_per_sm, s); else launch_mega_z<BVH, KIND, RT, 3>(P, blocks_per_sm, s); }
  else if constexpr (KIND != K_SIMPLE) launch_mega_z<BVH, KIND, RT, 1>(P, blocks_per_sm, s);
  else launch_mega_z<BVH, KIND, RT, 0>(P, blocks_per_sm, s);
}
template <int BVH, int KIND>
static void launch_mega_rt(const MegaParams& P, int blocks_per_sm, cudaStream_t s) {
  switch (P.rp.render_type) {
    case 0: launch_mega_t<BVH, KIND, 0>(P, blocks_per_sm, s); break;
    case 2: launch_mega_t<BVH, KIND, 2>(P, blocks_per_sm, s); break;
    default: launch_mega_t<BVH, KIND, 1>(P, blocks_per_sm, s); break;
  }
}
// blocks_per_sm[variant]: 0 = triangles/planes BVH2, 1 = triangles/planes BVH4, 2 = tori/boxes/ext BVH2, 3 = tori/boxes/ext BVH4
void launch_mega(const MegaParams& P, const int blocks_per_sm[4], cudaStream_t s) {
  if (!P.nslots) return;
  const bool b4 = P.rp.scene.bvh_kind == 4;
  const int v = (P.scene_kind == K_SIMPLE ? 0 : 2) + (b4 ? 1 : 0);
  switch (P.scene_kind) {
    case K_SIMPLE: if (b4) launch_mega_rt<4, K_SIMPLE>(P, blocks_per_sm[v], s); else launch_mega_rt<2, K_SIMPLE>(P, blocks_per_sm[v], s); break;
    case K_REF: if (b4) launch_mega_rt<4, K_REF>(P, blocks_per_sm[v], s); else launch_mega_rt<2, K_REF>(P, blocks_per_sm[v], s); break;
    default: if (b4) launch_mega_rt<4, K_EXT>(P, blocks_per_sm[v], s); else launch_mega_rt<2, K_EXT>(P, blocks_per_sm[v], s); break;
  }
}

// ------------------------------------------------------------------ segments (contract B10)
// accum[pix] += seg_0; += seg_1; ... in segment order, one thread per pixel; the sample count grows by spp.
// Pixel slots from `zone_pslot` on (the end zones of the queue) stored every sample's colour on its own: their segment sums
// are formed here, from +0 in sample order — the same additions as a lane makes for a whole segment.
__global__ void k_combine_segments(float4* accum, const uint32_t* __restrict__ pixel, uint32_t npix, const float4* __restrict__ seg_buf, uint32_t nseg, uint32_t spp,
                                   uint32_t zone_pslot, uint32_t seg_len) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  uint32_t pix = pixel[i];
  float4 a = accum[pix];
  if (i < zone_pslot) {
    for (uint32_t j = 0; j < nseg; j++) { float4 g = seg_buf[(size_t)i * nseg + j]; a.x += g.x; a.y += g.y; a.z += g.z; }
  } else {
    const float4* sp = seg_buf + (size_t)zone_pslot * nseg + (size_t)(i - zone_pslot) * spp;
    for (uint32_t j = 0; j < nseg; j++) {
      float sx = 0.0f, sy = 0.0f, sz = 0.0f;
      for (uint32_t k = j * seg_len; k < min((j + 1u) * seg_len, spp); k++) { float4 g = sp[k]; sx += g.x; sy += g.y; sz += g.z; }
      a.x += sx; a.y += sy; a.z += sz;
    }
  }
  a.w = __uint_as_float(__float_as_uint(a.w) + spp);
  accum[pix] = a;
}
void launch_combine_segments(float4* accum, const uint32_t* pixel, uint32_t npix, const float4* seg_buf, uint32_t nseg, uint32_t spp, uint32_t zone_pslot, uint32_t seg_len, cudaStream_t s) {
  if (!npix) return;
  k_combine_segments<<<(npix + 255) / 256, 256, 0, s>>>(accum, pixel, npix, seg_buf, nseg, spp, zone_pslot, seg_len);
}
// ---- strategy rounds: every pixel slot has its own sample count (0..33 for an adaptive round). The segments of all
// pixels are listed pixel by pixel (exclusive scan of ceil(spp / seg_len)); pixels without samples get no slot at all.
__global__ void k_seg_count(const uint32_t* __restrict__ slot_spp, uint32_t npix, uint32_t seg_len, uint32_t* __restrict__ cnt) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > npix) return;
  cnt[i] = i < npix ? (slot_spp[i] + seg_len - 1u) / seg_len : 0u;   // cnt[npix] = 0: the scan's last entry is the total
}
__global__ void k_seg_fill(const uint32_t* __restrict__ off, uint32_t npix, uint32_t* __restrict__ list, uint32_t shift) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  uint32_t a = off[i], b = off[i + 1];
  for (uint32_t j = 0; a + j < b; j++) list[a + j] = (i << shift) | j;
}
size_t seg_scan_bytes(uint32_t npix) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(npix + 1));
  return bytes;
}
void launch_build_segment_list(const uint32_t* slot_spp, uint32_t npix, uint32_t seg_len, uint32_t* seg_cnt, uint32_t* seg_off, void* scan_tmp, size_t scan_bytes, uint32_t* seg_list, uint32_t shift, cudaStream_t s) {
  if (!npix) return;
  k_seg_count<<<(npix + 1 + 255) / 256, 256, 0, s>>>(slot_spp, npix, seg_len, seg_cnt);
  cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, seg_cnt, seg_off, (int)(npix + 1), s);
  k_seg_fill<<<(npix + 255) / 256, 256, 0, s>>>(seg_off, npix, seg_list, shift);
}
// list_len != 0: the round was cut into short slots that stored one colour per sample (slot k owns entries k * list_len ..); the
// segment sums (seg_len samples from +0, in sample order) are formed here — the same additions as a lane makes for a whole segment.
__global__ void k_combine_segment_list(float4* accum, const uint32_t* __restrict__ pixel, uint32_t npix, const float4* __restrict__ seg_buf, const uint32_t* __restrict__ off, const uint32_t* __restrict__ slot_spp,
                                       uint32_t list_len, uint32_t seg_len) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  uint32_t a = off[i], b = off[i + 1];
  if (a == b) return;
  uint32_t pix = pixel[i];
  float4 acc = accum[pix];
  if (list_len) {
    const float4* sp = seg_buf + (size_t)a * list_len;
    const uint32_t spp = slot_spp[i];
    for (uint32_t k0 = 0; k0 < spp; k0 += seg_len) {
      float sx = 0.0f, sy = 0.0f, sz = 0.0f;
      for (uint32_t k = k0; k < min(k0 + seg_len, spp); k++) { float4 g = sp[k]; sx += g.x; sy += g.y; sz += g.z; }
      acc.x += sx; acc.y += sy; acc.z += sz;
    }
  } else
  for (uint32_t k = a; k < b; k++) { float4 g = seg_buf[k]; acc.x += g.x; acc.y += g.y; acc.z += g.z; }
  acc.w = __uint_as_float(__float_as_uint(acc.w) + slot_spp[i]);
  accum[pix] = acc;
}
void launch_combine_segment_list(float4* accum, const uint32_t* pixel, uint32_t npix, const float4* seg_buf, const uint32_t* seg_off, const uint32_t* slot_spp, uint32_t list_len, uint32_t seg_len, cudaStream_t s) {
  if (!npix) return;
  k_combine_segment_list<<<(npix + 255) / 256, 256, 0, s>>>(accum, pixel, npix, seg_buf, seg_off, slot_spp, list_len, seg_len);
}
// wavefront engine: the samples of pass `pass` (= segment index) of every slot; any_left counts slots that still have samples
__global__ void k_segment_pass_spp(const uint32_t* __restrict__ slot_spp, uint32_t npix, uint32_t seg_len, uint32_t pass, uint32_t* __restrict__ pass_spp, uint32_t* any_left) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  uint32_t spp = slot_spp[i], b = min(pass * seg_len, spp), m = min(seg_len, spp - b);
  pass_spp[i] = m;
  if (m && any_left) atomicAdd(any_left, 1u);
}
void launch_segment_pass_spp(const uint32_t* slot_spp, uint32_t npix, uint32_t seg_len, uint32_t pass, uint32_t* pass_spp, uint32_t* any_left, cudaStream_t s) {
  if (!npix) return;
  k_segment_pass_spp<<<(npix + 255) / 256, 256, 0, s>>>(slot_spp, npix, seg_len, pass, pass_spp, any_left);
}
// wavefront engine: one segment was accumulated sample by sample into seg_acc (rgb sum from +0, count): add it
__global__ void k_add_segment(float4* accum, const uint32_t* __restrict__ pixel, uint32_t npix, const float4* __restrict__ seg_acc) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  uint32_t pix = pixel[i];
  float4 a = accum[pix], g = seg_acc[pix];
  a.x += g.x; a.y += g.y; a.z += g.z;
  a.w = __uint_as_float(__float_as_uint(a.w) + __float_as_uint(g.w));
  accum[pix] = a;
}
void launch_add_segment(float4* accum, const uint32_t* pixel, uint32_t npix, const float4* seg_acc, cudaStream_t s) {
  if (!npix) return;
  k_add_segment<<<(npix + 255) / 256, 256, 0, s>>>(accum, pixel, npix, seg_acc);
}
__global__ void k_clear_pixels(float4* buf, const uint32_t* __restrict__ pixel, uint32_t npix) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npix) buf[pixel[i]] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}
void launch_clear_pixels(float4* buf, const uint32_t* pixel, uint32_t npix, cudaStream_t s) {
  if (!npix) return;
  k_clear_pixels<<<(npix + 255) / 256, 256, 0, s>>>(buf, pixel, npix);
}

// ------------------------------------------------------------------ resolve (render_target.rs:59-64)
WPT_DEV uint32_t to_u8(float v) {   // `( x.min(1.0).max(0.0) * 255.0 ) as u8` — trunc, saturating, NaN -> 0
  float c = fmaxf(fminf(v, 1.0f), 0.0f) * 255.0f;
  if (!(c > 0.0f)) return 0u;
  if (c >= 255.0f) return 255u;
  return (uint32_t)c;
}
__global__ void k_resolve_rgba(const float4* __restrict__ accum, uint32_t* __restrict__ rgba, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = accum[i];
  uint32_t cnt = __float_as_uint(a.w);
  uint32_t out = 0xFF000000u;
  if (cnt) {
    float c = (float)cnt;
    out |= to_u8(a.x / c) | (to_u8(a.y / c) << 8) | (to_u8(a.z / c) << 16);
  }
  rgba[i] = out;
}
void launch_resolve_rgba(const float4* accum, uint8_t* rgba, uint32_t n, cudaStream_t s) {
  if (!n) return;
  k_resolve_rgba<<<(n + 255) / 256, 256, 0, s>>>(accum, reinterpret_cast<uint32_t*>(rgba), n);
}

// ------------------------------------------------------------------ photons (tracer.rs:126-152)
// One thread per photon shot k (stream (k, 0, STREAM_PHOTON)): light pick, point on the light,
// uniform hemisphere direction by rejection (rng.rs:50-68), one Scene::trace, store on a
// diffuse hit. `meta[i]` = node visits | (stored ? 1<<31 : 0) for the host-side cut. The batch buffers are
// zeroed before the launch and every slot is written by exactly one rank, so a sum-allreduce of the raw
// 32-bit words merges the ranks' shots bit for bit (x + 0 = x).
WPT_DEV F3 next_hemisphere(Rng& rng, F3 normal) {
  float x, y, z;
  for (;;) {
    x = rng.f32() * 2.0f - 1.0f;
    y = rng.f32() * 2.0f - 1.0f;
    z = rng.f32() * 2.0f - 1.0f;
    float ls = x * x + y * y + z * z;
    if (!(ls > 1.0f)) break;
  }
  F3 v = normalize(f3(x, y, z));
  if (dot(v, normal) < 0.0f) return -v;
  return v;
}
__global__ void __launch_bounds__(128) k_photon_emit(RenderParams rp, unsigned long long shot0, uint32_t n, uint32_t rank, uint32_t world, uint32_t* meta, uint32_t* rec_light, float4* rec_loc_w) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || i % world != rank) return;   // multi-GPU: rank r emits the shots r, r + world, ... of the batch
  unsigned long long k = shot0 + i;
  Rng rng; rng.s = stream_seed((uint32_t)k, 0u, STREAM_PHOTON, rp.base_seed);
  uint32_t light_id = rng.range(0, rp.scene.num_lights);
  F3 pl, ln, inten; float area; uint32_t lsid;
  pick_random(rp.scene, light_id, rng, &pl, &ln, &inten, &area, &lsid);
  F3 dir = next_hemisphere(rng, ln);
  Ray ray = make_ray(pl + dir * WPT_EPSILON, dir);
  GHit g = trace_g(rp.scene, ray);
  bool stored = false;
  if (g.id >= 0) {
    float t; F3 n; uint32_t mat;
    if (shape_trace_full<K_EXT>(rp.scene.shapes, (uint32_t)g.id, ray, &t, &n, &mat)) {
      float4 mc = __ldg(&rp.scene.mats[mat].c);
      if (mc.w == (float)MAT_DIFFUSE || mc.w == (float)MAT_DIFFUSE_TEX) {   // hit.mat.is_diffuse(), tracer.rs:144
        F3 hp = (ray.o + t * ray.d) + n * WPT_EPSILON;
        float w = dot(ln, dir) * fmaxf(fmaxf(inten.x, inten.y), inten.z);
        rec_loc_w[i] = make_float4(hp.x, hp.y, hp.z, w); rec_light[i] = light_id;   // dense: the record of shot i lives in slot i
        stored = true;
      }
    }
  }
  meta[i] = g.visits | (stored ? 0x80000000u : 0u);
}
void launch_photon_emit(const RenderParams& rp, unsigned long long shot0, uint32_t n, uint32_t rank, uint32_t world, uint32_t* meta, uint32_t* rec_light, float4* rec_loc_w, cudaStream_t s) {
  if (!n) return;
  k_photon_emit<<<(n + 127) / 128, 128, 0, s>>>(rp, shot0, n, rank, world ? world : 1u, meta, rec_light, rec_loc_w);
}

// Batch compaction on the device: off[i] = photons stored by the shots before shot i of this batch (exclusive scan of the stored
// flags). The photon set ends with the shot that stores photon number `target`: shot i belongs to it iff off[i] < remaining.
// res[0] += photons taken, res[1] += shots taken (= the cut), res[2] += node visits and res[3] += shots of this rank's own shots.
__global__ void k_photon_flags(const uint32_t* __restrict__ meta, uint32_t n, uint32_t* __restrict__ flag) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) flag[i] = i < n ? meta[i] >> 31 : 0u;
}
__global__ void k_photon_take(const uint32_t* __restrict__ meta, const uint32_t* __restrict__ light, const float4* __restrict__ lw, const uint32_t* __restrict__ off, uint32_t n,
                              uint32_t remaining, uint32_t base, uint32_t rank, uint32_t world, float4* __restrict__ out_lw, uint2* __restrict__ out_ls, unsigned long long* res) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long taken = 0, shots = 0, visits = 0, own = 0;
  if (i < n) {
    const uint32_t m = meta[i], o = off[i];
    if (o < remaining) {
      shots = 1;
      if (i % world == rank) { visits = m & 0x7FFFFFFFu; own = 1; }
      if (m >> 31) { out_lw[base + o] = lw[i]; out_ls[base + o] = make_uint2(light[i], i); taken = 1; }
    }
  }
  taken = warp_sum_u64(taken); shots = warp_sum_u64(shots); visits = warp_sum_u64(visits); own = warp_sum_u64(own);
  if ((threadIdx.x & 31) == 0) {
    if (taken) atomicAdd(&res[0], taken);
    if (shots) atomicAdd(&res[1], shots);
    if (visits) atomicAdd(&res[2], visits);
    if (own) atomicAdd(&res[3], own);
  }
}
void launch_photon_compact(const uint32_t* meta, const uint32_t* light, const float4* lw, uint32_t n, uint32_t remaining, uint32_t base, uint32_t rank, uint32_t world,
                           uint32_t* flag, uint32_t* off, void* scan_tmp, size_t scan_bytes, float4* out_lw, uint2* out_ls, unsigned long long* res, cudaStream_t s) {
  if (!n) return;
  k_photon_flags<<<(n + 1 + 255) / 256, 256, 0, s>>>(meta, n, flag);
  cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flag, off, (int)(n + 1), s);
  k_photon_take<<<(n + 255) / 256, 256, 0, s>>>(meta, light, lw, off, n, remaining, base, rank, world ? world : 1u, out_lw, out_ls, res);
}

// ---- octree (photon_tree.rs). Cells are split level by level; a cell is split iff it finally
// holds more than 1024 photons (photon_tree.rs:29,179), so the topology does not depend on
// insertion order. node_of[p] = the deepest existing cell containing photon p.
WPT_DEV uint32_t tree_descend_child(Cell& b, F3 v) { return octree_child(b, v); }
__global__ void k_octree_assign(const float4* __restrict__ loc_w, uint32_t n, uint32_t* node_of, const uint32_t* __restrict__ child_base, uint32_t* count) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t node = node_of[p];
  uint32_t cb = child_base[node];
  if (cb == 0xFFFFFFFFu) return;   // still a leaf
  // find the bounds of `node` by walking down from the root, then step into the child
  F3 v = xyz(loc_w[p]);
  Cell b = {-1024.0f, -1024.0f, -1024.0f, 1024.0f, 1024.0f, 1024.0f};
  uint32_t cur = 0;
  while (cur != node) cur = child_base[cur] + octree_child(b, v);
  uint32_t child = cb + octree_child(b, v);
  node_of[p] = child;
  atomicAdd(&count[child], 1u);
}
void launch_octree_assign(const float4* loc_w, uint32_t n, uint32_t* node_of, const uint32_t* child_base, uint32_t* count, cudaStream_t s) {
  if (!n) return;
  k_octree_assign<<<(n + 255) / 256, 256, 0, s>>>(loc_w, n, node_of, child_base, count);
}
// per-node, per-light weight sums in 2^-40 fixed point (order independent): every
// photon adds its weight to each cell on its root-to-leaf path (photon_tree.rs:168,176)
WPT_DEV unsigned long long weight_fx(float w) {
  double sc = (double)w * 1099511627776.0;
  if (!(sc > 0.0)) return 0ull;
  return (unsigned long long)__double2ll_rn(sc);
}
__global__ void k_octree_bins(const float4* __restrict__ loc_w, const uint2* __restrict__ light_shot, uint32_t n, const uint32_t* __restrict__ child_base, unsigned long long* fx, uint32_t num_lights) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float4 lw = loc_w[p];
  F3 v = xyz(lw);
  unsigned long long w = weight_fx(lw.w);
  uint32_t light = light_shot[p].x;
  Cell b = {-1024.0f, -1024.0f, -1024.0f, 1024.0f, 1024.0f, 1024.0f};
  uint32_t cur = 0;
  for (;;) {
    atomicAdd(&fx[(size_t)cur * num_lights + light], w);
    uint32_t cb = child_base[cur];
    if (cb == 0xFFFFFFFFu) break;
    cur = cb + octree_child(b, v);
  }
}
void launch_octree_bins(const float4* loc_w, const uint2* light_shot, uint32_t n, const uint32_t* child_base, unsigned long long* fx, uint32_t num_lights, cudaStream_t s) {
  if (!n) return;
  k_octree_bins<<<(n + 255) / 256, 256, 0, s>>>(loc_w, light_shot, n, child_base, fx, num_lights);
}
// EmpiricalPDF::recheck_cdf (empirical_pdf.rs:79-93): bins start at 1.0 (:24)
__global__ void k_octree_cdf(const unsigned long long* __restrict__ fx, float* bins, float* cum, uint32_t num_nodes, uint32_t num_lights) {
  uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= num_nodes) return;
  const unsigned long long* f = fx + (size_t)node * num_lights;
  float* b = bins + (size_t)node * num_lights;
  float* c = cum + (size_t)node * num_lights;
  float bin_sum = 0.0f;
  for (uint32_t i = 0; i < num_lights; i++) {
    float v = 1.0f + (float)(__ull2double_rn(f[i]) * (1.0 / 1099511627776.0));
    b[i] = v;
    bin_sum += v;
  }
  c[0] = 0.0f;
  for (uint32_t i = 1; i < num_lights; i++) c[i] = c[i - 1] + b[i - 1] / bin_sum;
}
void launch_octree_cdf(const unsigned long long* fx, float* bins, float* cum, uint32_t num_nodes, uint32_t num_lights, cudaStream_t s) {
  if (!num_nodes) return;
  k_octree_cdf<<<(num_nodes + 127) / 128, 128, 0, s>>>(fx, bins, cum, num_nodes, num_lights);
}
__global__ void k_photon_sample_batch(DPhotonTree t, const float* __restrict__ pts, const uint32_t* __restrict__ seeds, uint64_t n, uint32_t* light, float* pdf) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Rng rng; rng.s = seeds[i];
  uint32_t l; float p;
  photon_sample(t, rng, f3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]), &l, &p);
  light[i] = l; pdf[i] = p;
}
void launch_photon_sample_batch(const DPhotonTree& t, const float* pts, const uint32_t* seeds, uint64_t n, uint32_t* light, float* pdf, cudaStream_t s) {
  if (!n) return;
  k_photon_sample_batch<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(t, pts, seeds, n, light, pdf);
}

// ------------------------------------------------------------------ adaptive sampling (sampling_strategy.rs:122-182)
WPT_DEV F3 read_clamped(const float4* __restrict__ accum, uint32_t W, int x, int y) {   // render_target.rs:74-77
  float4 a = accum[(size_t)y * W + x];
  float c = (float)__float_as_uint(a.w);
  F3 v = f3(a.x / c, a.y / c, a.z / c);
  return f3(fminf(fmaxf(v.x, 0.0f), 1.0f), fminf(fmaxf(v.y, 0.0f), 1.0f), fminf(fmaxf(v.z, 0.0f), 1.0f));
}
template <int K> WPT_DEV F3 gaussian(const float4* __restrict__ accum, uint32_t W, uint32_t H, int x, int y) {   // render_target.rs:88-138
  const float g3[9] = {1, 2, 1, 2, 4, 2, 1, 2, 1};
  const float g5[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};
  float sum = 0.0f; F3 acc = f3(0, 0, 0);
  for (int vy = 0; vy < K; vy++)
    for (int vx = 0; vx < K; vx++) {
      int px = x + vx - K / 2, py = y + vy - K / 2;
      float m = K == 3 ? g3[vy * 3 + vx] : g5[vy * 5 + vx];
      if (px < 0 || py < 0 || px >= (int)W || py >= (int)H) { acc = acc + f3(0, 0, 0); sum += 0.0f; }
      else { acc = acc + m * read_clamped(accum, W, px, py); sum += m; }
    }
  return acc / sum;
}
// error per pixel of the region + {sum in 2^-40 fixed point, min, max}; stats[0]=sum (u64),
// stats[1]=min bits, stats[2]=max bits (errors are >= 0, so uint order == float order)
// One block = 32 x 8 pixels of the region; the clamped means of its 36 x 12 neighbourhood (two pixels of halo for the 5 x 5
// Gaussian) are computed once into shared memory — 1.7 loads and divisions per pixel instead of 35. Same values, same order of
// the weighted sums as `gaussian` above (an out-of-frame neighbour contributes 0 * 0 = +0 and weight 0, as `acc + 0`, `sum + 0`
// there); the Gaussians read across the region boundary, inside the frame (quirk q9).
#define EM_BX 32
#define EM_BY 8
__global__ void __launch_bounds__(EM_BX * EM_BY) k_error_map(const float4* __restrict__ accum, uint32_t W, uint32_t H, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, float* mse, unsigned long long* stats, const unsigned long long* gate) {
  if (gate && *gate != AD_ERR) return;   // device-driven rounds: only when this step opens a new adaptive round
  constexpr int TW = EM_BX + 4, TH = EM_BY + 4;
  __shared__ float tr[TH][TW], tg[TH][TW], tb[TH][TW], tin[TH][TW];
  const int x0 = (int)(rx + blockIdx.x * EM_BX) - 2, y0 = (int)(ry + blockIdx.y * EM_BY) - 2;
  for (int k = threadIdx.y * EM_BX + threadIdx.x; k < TW * TH; k += EM_BX * EM_BY) {
    const int ty = k / TW, tx = k - ty * TW, px = x0 + tx, py = y0 + ty;
    const bool in = px >= 0 && py >= 0 && px < (int)W && py < (int)H;
    F3 v = f3(0, 0, 0);
    if (in) v = read_clamped(accum, W, px, py);
    tr[ty][tx] = v.x; tg[ty][tx] = v.y; tb[ty][tx] = v.z; tin[ty][tx] = in ? 1.0f : 0.0f;
  }
  __syncthreads();
  const uint32_t lx = blockIdx.x * EM_BX + threadIdx.x, ly = blockIdx.y * EM_BY + threadIdx.y;
  unsigned long long fx = 0; uint32_t mn = 0x7F800000u, mx = 0u;
  if (lx < rw && ly < rh) {
    const int cx = threadIdx.x + 2, cy = threadIdx.y + 2;
    const float g3[9] = {1, 2, 1, 2, 4, 2, 1, 2, 1};
    const float g5[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};
    const F3 v0 = f3(tr[cy][cx], tg[cy][cx], tb[cy][cx]);
    float s1 = 0.0f; F3 a1 = f3(0, 0, 0);
#pragma unroll
    for (int vy = 0; vy < 3; vy++)
#pragma unroll
      for (int vx = 0; vx < 3; vx++) {
        const int yy = cy + vy - 1, xx = cx + vx - 1;
        const float m = g3[vy * 3 + vx] * tin[yy][xx];
        a1 = a1 + m * f3(tr[yy][xx], tg[yy][xx], tb[yy][xx]); s1 += m;
      }
    float s2 = 0.0f; F3 a2 = f3(0, 0, 0);
#pragma unroll
    for (int vy = 0; vy < 5; vy++)
#pragma unroll
      for (int vx = 0; vx < 5; vx++) {
        const int yy = cy + vy - 2, xx = cx + vx - 2;
        const float m = g5[vy * 5 + vx] * tin[yy][xx];
        a2 = a2 + m * f3(tr[yy][xx], tg[yy][xx], tb[yy][xx]); s2 += m;
      }
    const F3 v1 = a1 / s1, v2 = a2 / s2;
    const F3 d1 = v0 - v1, d2 = v0 - v2;
    const float e = fmaxf(dot(d1, d1), dot(d2, d2));
    mse[(size_t)ly * rw + lx] = e;
    fx = weight_fx(e); mn = __float_as_uint(e); mx = mn;
  }
  fx = warp_sum_u64(fx);
  for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_down_sync(0xFFFFFFFFu, mn, o)); mx = max(mx, __shfl_down_sync(0xFFFFFFFFu, mx, o)); }
  if (threadIdx.x == 0) {
    atomicAdd(&stats[0], fx);
    atomicMin(reinterpret_cast<unsigned int*>(&stats[1]), mn);
    atomicMax(reinterpret_cast<unsigned int*>(&stats[2]), mx);
  }
}
void launch_error_map(const float4* accum, uint32_t W, uint32_t H, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, float* mse, unsigned long long* stats, const unsigned long long* gate, cudaStream_t s) {
  if (!rw || !rh) return;
  k_error_map<<<dim3((rw + EM_BX - 1) / EM_BX, (rh + EM_BY - 1) / EM_BY), dim3(EM_BX, EM_BY), 0, s>>>(accum, W, H, rx, ry, rw, rh, mse, stats, gate);
}
WPT_DEV uint32_t sampling_rgba(F3 v) { return 0xFF000000u | to_u8(v.x) | (to_u8(v.y) << 8) | (to_u8(v.z) << 16); }
// error -> samples this round (1..33) + the sampling-density view (sampling_strategy.rs:154-174)
// stats (device): [0] = sum of the errors in 2^-40 fixed point, [1] = min bits, [2] = max bits (k_error_map), [3] = total of the
// samples this round allocates (written here): the host reads one word per round instead of synchronising twice
__global__ void k_adaptive_spp(const float* __restrict__ mse, uint32_t n, unsigned long long* stats, uint32_t* round_left, uint32_t* round_spp, uint32_t* sampling_rgba8, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, const unsigned long long* gate) {
  if (gate && *gate != AD_ERR) return;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const float mn = __uint_as_float((uint32_t)stats[1]), mx = __uint_as_float((uint32_t)stats[2]);
  const float avg = (float)(((double)stats[0] * (1.0 / 1099511627776.0)) / (double)n);
  unsigned long long mine = 0;
  if (i < n) {
    float e = mse[i];
    float sc = (e < avg) ? 0.5f * ((e - mn) / (avg - mn)) : 0.5f + 0.5f * ((e - avg) / (mx - avg));
    sc = fmaxf(fminf(sc, 1.0f), 0.0f);
    float c = ceilf(1.0f + sc * 32.0f);
    uint32_t spp = c > 0.0f ? (uint32_t)c : 0u;
    round_left[i] = spp; if (round_spp) round_spp[i] = spp; mine = spp;
    F3 col;
    if (mn == mx) col = f3(0, 0, 0);
    else if (sc < 0.5f) col = f3(0.0f, 1.0f, 0.0f) * (1.0f - 2.0f * sc) + f3(0.0f, 0.0f, 1.0f) * 2.0f * sc;   // mix_color, :224-230
    else col = f3(0.0f, 0.0f, 1.0f) * (1.0f - 2.0f * (sc - 0.5f)) + f3(1.0f, 0.0f, 0.0f) * 2.0f * (sc - 0.5f);
    sampling_rgba8[(size_t)(ry + i / rw) * W + rx + i % rw] = sampling_rgba(col);
  }
  mine = warp_sum_u64(mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&stats[3], mine);
}
void launch_adaptive_spp(const float* mse, uint32_t n, unsigned long long* stats, uint32_t* round_left, uint32_t* round_spp, uint8_t* sampling_rgba8, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, const unsigned long long* gate, cudaStream_t s) {
  if (!n) return;
  k_adaptive_spp<<<(n + 255) / 256, 256, 0, s>>>(mse, n, stats, round_left, round_spp, reinterpret_cast<uint32_t*>(sampling_rgba8), W, rx, ry, rw, gate);
}
// fill a region of the sampling view / a u32 array
__global__ void k_fill_region_rgba(uint32_t* rgba, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, uint32_t value) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rw * rh) return;
  rgba[(size_t)(ry + i / rw) * W + rx + i % rw] = value;
}
void launch_fill_region_rgba(uint8_t* rgba, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh, uint32_t value, cudaStream_t s) {
  if (!rw || !rh) return;
  k_fill_region_rgba<<<(rw * rh + 255) / 256, 256, 0, s>>>(reinterpret_cast<uint32_t*>(rgba), W, rx, ry, rw, rh, value);
}
__global__ void k_fill_u32(uint32_t* a, uint32_t n, uint32_t v) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
void launch_fill_u32(uint32_t* a, uint32_t n, uint32_t v, cudaStream_t s) {
  if (!n) return;
  k_fill_u32<<<(n + 255) / 256, 256, 0, s>>>(a, n, v);
}
// total of round_left (u64) — decides whether a budget cut is needed
__global__ void k_sum_u32(const uint32_t* __restrict__ a, uint32_t n, unsigned long long* out) {
  unsigned long long v = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += a[i];
  v = warp_sum_u64(v);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}
void launch_sum_u32(const uint32_t* a, uint32_t n, unsigned long long* out, cudaStream_t s) {
  if (!n) return;
  int grid = (int)min((n + 255u) / 256u, 1184u);
  k_sum_u32<<<grid, 256, 0, s>>>(a, n, out);
}
// Budget cut in the reference's pop order (LIFO over raster pushes = from the last pixel
// backwards): take[i] = clamp(budget - sum_{j>i} left[j], 0, left[i]). Two passes: per-block
// totals, then a block-local reverse scan with the suffix of the later blocks.
#define CUT_BLOCK 1024
__global__ void k_cut_block_totals(const uint32_t* __restrict__ left, uint32_t n, unsigned long long* block_tot) {
  __shared__ unsigned long long sh[CUT_BLOCK / 32];
  uint32_t i = blockIdx.x * CUT_BLOCK + threadIdx.x;
  unsigned long long v = i < n ? left[i] : 0;
  v = warp_sum_u64(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { unsigned long long t = 0; for (int k = 0; k < CUT_BLOCK / 32; k++) t += sh[k]; block_tot[blockIdx.x] = t; }
}
__global__ void k_cut_apply(const uint32_t* __restrict__ left, uint32_t n, const unsigned long long* __restrict__ block_suffix, unsigned long long budget, const unsigned long long* budget_dev, uint32_t* take) {
  if (budget_dev) budget = *budget_dev;   // device-driven rounds: the room left in the call's budget
  // block_suffix[b] = sum of left[] over all blocks after b
  __shared__ unsigned long long sh[CUT_BLOCK];
  uint32_t i = blockIdx.x * CUT_BLOCK + threadIdx.x;
  sh[threadIdx.x] = i < n ? left[i] : 0;
  __syncthreads();
  // reverse inclusive scan in shared memory (Hillis-Steele on the reversed index)
  for (int off = 1; off < CUT_BLOCK; off <<= 1) {
    unsigned long long add = (threadIdx.x + off < CUT_BLOCK) ? sh[threadIdx.x + off] : 0;
    __syncthreads();
    sh[threadIdx.x] += add;
    __syncthreads();
  }
  if (i < n) {
    unsigned long long own = left[i];
    unsigned long long after = block_suffix[blockIdx.x] + (sh[threadIdx.x] - own);   // samples queued behind this pixel
    unsigned long long room = budget > after ? budget - after : 0;
    take[i] = (uint32_t)(own < room ? own : room);
  }
}
void launch_cut(const uint32_t* left, uint32_t n, unsigned long long* block_tot, const unsigned long long* block_suffix, unsigned long long budget, uint32_t* take, int pass, cudaStream_t s) {
  if (!n) return;
  uint32_t blocks = (n + CUT_BLOCK - 1) / CUT_BLOCK;
  if (pass == 0) k_cut_block_totals<<<blocks, CUT_BLOCK, 0, s>>>(left, n, block_tot);
  else k_cut_apply<<<blocks, CUT_BLOCK, 0, s>>>(left, n, block_suffix, budget, nullptr, take);
}
// ---- device-driven adaptive rounds (Context::run_adaptive): the state of AdaptiveSamplingStrategy (sampling_strategy.rs:77-220)
// lives on the device, the host only enqueues steps. st[0..2] error stats, [3] total of the round being opened, [4] samples left
// in the current round, [5] ticks used by this call, [6] the call's budget, [7] first queue issued, [8] ticks this step takes,
// [9] mode of this step, [10] steps that rendered, [11] room left in the budget.
__global__ void k_ad_setup(unsigned long long* st, unsigned long long budget) { st[5] = 0; st[6] = budget; st[10] = 0; }
__global__ void k_ad_begin(unsigned long long* st) {
  unsigned long long mode;
  if (st[5] >= st[6]) mode = AD_IDLE;                    // budget spent: the remaining enqueued steps do nothing
  else if (st[4] != 0) mode = AD_CONT;                   // a round cut by an earlier budget continues
  else if (!st[7]) { mode = AD_FIRST; st[7] = 1; }       // first queue: 4 samples per pixel (sampling_strategy.rs:197-203)
  else { mode = AD_ERR; st[0] = 0; st[1] = 0x7F800000ull; st[2] = 0; }
  st[3] = 0; st[9] = mode;
}
__global__ void k_ad_first(unsigned long long* st, uint32_t* round_left, uint32_t* round_spp, uint32_t n) {
  if (st[9] != AD_FIRST) return;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { round_left[i] = 4u; round_spp[i] = 4u; }
  if (i == 0) st[3] = 4ull * n;
}
__global__ void k_ad_total(unsigned long long* st) {
  if (st[9] == AD_FIRST || st[9] == AD_ERR) st[4] = st[3];
  const unsigned long long room = st[9] == AD_IDLE ? 0ull : st[6] - st[5];
  st[11] = room;
  st[8] = st[4] < room ? st[4] : room;
}
__global__ void k_ad_end(unsigned long long* st) {
  const unsigned long long taken = st[8];
  st[4] -= taken; st[5] += taken;
  if (taken) st[10] += 1;
}
// block_suffix[b] = samples queued in the blocks after b. One block of 1024 threads: every thread sums its run of consecutive
// entries, a shared-memory suffix scan gives what lies behind the run, then the run is written back to front (integer sums: exact
// in any order). A single thread walking the few thousand entries took 93 us per round (ncu launch list of the target frame).
__global__ void __launch_bounds__(1024) k_cut_suffix(const unsigned long long* __restrict__ block_tot, uint32_t blocks, unsigned long long* block_suffix) {
  __shared__ unsigned long long sh[1024];
  const uint32_t t = threadIdx.x, per = (blocks + 1023u) / 1024u;
  const uint32_t lo = min(t * per, blocks), hi = min(lo + per, blocks);
  unsigned long long sum = 0;
  for (uint32_t i = lo; i < hi; i++) sum += block_tot[i];
  sh[t] = sum;
  __syncthreads();
  for (uint32_t o = 1; o < 1024u; o <<= 1) {
    const unsigned long long v = t + o < 1024u ? sh[t + o] : 0ull;
    __syncthreads();
    sh[t] += v;
    __syncthreads();
  }
  unsigned long long run = sh[t] - sum;   // the entries behind this thread's run
  for (uint32_t i = hi; i-- > lo;) { block_suffix[i] = run; run += block_tot[i]; }
}
void launch_ad_setup(unsigned long long* st, unsigned long long budget, cudaStream_t s) { k_ad_setup<<<1, 1, 0, s>>>(st, budget); }
void launch_ad_begin(unsigned long long* st, cudaStream_t s) { k_ad_begin<<<1, 1, 0, s>>>(st); }
void launch_ad_first(unsigned long long* st, uint32_t* round_left, uint32_t* round_spp, uint32_t n, cudaStream_t s) { if (n) k_ad_first<<<(n + 255) / 256, 256, 0, s>>>(st, round_left, round_spp, n); }
void launch_ad_total(unsigned long long* st, cudaStream_t s) { k_ad_total<<<1, 1, 0, s>>>(st); }
void launch_ad_end(unsigned long long* st, cudaStream_t s) { k_ad_end<<<1, 1, 0, s>>>(st); }
void launch_cut_device(const uint32_t* left, uint32_t n, unsigned long long* block_tot, unsigned long long* block_suffix, const unsigned long long* room_dev, uint32_t* take, cudaStream_t s) {
  if (!n) return;
  uint32_t blocks = (n + CUT_BLOCK - 1) / CUT_BLOCK;
  k_cut_block_totals<<<blocks, CUT_BLOCK, 0, s>>>(left, n, block_tot);
  k_cut_suffix<<<1, 1024, 0, s>>>(block_tot, blocks, block_suffix);
  k_cut_apply<<<blocks, CUT_BLOCK, 0, s>>>(left, n, block_suffix, 0ull, room_dev, take);
}
// slot spp from the region-indexed take[]; round_left -= take for this session's rows only is
// done by the caller on the region array (all ranks hold the same region arrays)
__global__ void k_gather_slot_spp(const uint32_t* __restrict__ take, const uint32_t* __restrict__ pixel, uint32_t nslots, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t* slot_spp) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nslots) return;
  uint32_t pix = pixel[i];
  uint32_t y = pix / W, x = pix - y * W;
  slot_spp[i] = take[(y - ry) * rw + (x - rx)];
}
void launch_gather_slot_spp(const uint32_t* take, const uint32_t* pixel, uint32_t nslots, uint32_t W, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t* slot_spp, cudaStream_t s) {
  if (!nslots) return;
  k_gather_slot_spp<<<(nslots + 255) / 256, 256, 0, s>>>(take, pixel, nslots, W, rx, ry, rw, slot_spp);
}
__global__ void k_sub_u32(uint32_t* a, const uint32_t* __restrict__ b, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] -= b[i];
}
void launch_sub_u32(uint32_t* a, const uint32_t* b, uint32_t n, cudaStream_t s) {
  if (!n) return;
  k_sub_u32<<<(n + 255) / 256, 256, 0, s>>>(a, b, n);
}
// Random strategy (sampling_strategy.rs:56-59) in mode B: tick t picks its pixel from stream
// (t, t>>32, STREAM_PIXEL); take[pixel] counts the ticks of this call.
__global__ void k_random_ticks(unsigned long long t0, unsigned long long n, uint32_t seed, uint32_t rw, uint32_t rh, uint32_t* take) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long t = t0 + i;
  Rng rng; rng.s = stream_seed((uint32_t)t, (uint32_t)(t >> 32), STREAM_PIXEL, seed);
  uint32_t x = rng.range(0, rw);
  uint32_t y = rng.range(0, rh);
  atomicAdd(&take[y * rw + x], 1u);
}
void launch_random_ticks(unsigned long long t0, unsigned long long n, uint32_t seed, uint32_t rw, uint32_t rh, uint32_t* take, cudaStream_t s) {
  if (!n) return;
  k_random_ticks<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t0, n, seed, rw, rh, take);
}

// ------------------------------------------------------------------ probes
__global__ void k_primary_probe(RenderParams rp, int32_t* ids, uint32_t* visits, float* dist) {
  uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= rp.W * rp.H) return;
  Rng rng; rng.s = stream_seed(pix, 0, STREAM_PATH, rp.base_seed);
  float j1 = rng.f32();
  float j2 = rng.f32();
  uint32_t py = pix / rp.W, px = pix - py * rp.W;
  Ray ray = camera_ray(rp.cam, px, py, j1, j2);
  GHit g = trace_g(rp.scene, ray);
  if (ids) ids[pix] = g.id;
  if (visits) visits[pix] = g.visits;
  if (dist) dist[pix] = g.id >= 0 ? g.t : WPT_INF;
}
void launch_primary_probe(const RenderParams& rp, int32_t* ids, uint32_t* visits, float* dist, cudaStream_t s) {
  uint32_t n = rp.W * rp.H;
  k_primary_probe<<<(n + 127) / 128, 128, 0, s>>>(rp, ids, visits, dist);
}

__global__ void k_trace_batch(RenderParams rp, const float* __restrict__ o, const float* __restrict__ d, uint64_t n, int32_t* ids, float* dist, uint32_t* visits, float* normals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Ray ray = make_ray(f3(o[i * 3], o[i * 3 + 1], o[i * 3 + 2]), f3(d[i * 3], d[i * 3 + 1], d[i * 3 + 2]));
  GHit g = trace_g(rp.scene, ray);
  ids[i] = g.id;
  dist[i] = g.id >= 0 ? g.t : WPT_INF;
  visits[i] = g.visits;
  if (normals) {
    float t; F3 nn = f3(0, 0, 0); uint32_t mat;
    bool ok = g.id >= 0 && shape_trace_full<K_EXT>(rp.scene.shapes, (uint32_t)g.id, ray, &t, &nn, &mat);
    normals[i * 3] = ok ? nn.x : 0.0f; normals[i * 3 + 1] = ok ? nn.y : 0.0f; normals[i * 3 + 2] = ok ? nn.z : 0.0f;
  }
}
void launch_trace_batch(const RenderParams& rp, const float* o, const float* d, uint64_t n, int32_t* ids, float* dist, uint32_t* visits, float* normals, cudaStream_t s) {
  if (!n) return;
  k_trace_batch<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(rp, o, d, n, ids, dist, visits, normals);
}

}  // namespace wpt
