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

namespace wpt {

static int g_sm_count = 0;
int device_sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

#define TRACE_THREADS 128
#define SHADE_THREADS 128

// ------------------------------------------------------------------ helpers
WPT_DEV unsigned long long warp_sum_u64(unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
  return v;
}

// tracer.rs:176-191 — camera ray through pixel (x,y) with jitter (j1,j2)
WPT_DEV Ray camera_ray(const DCamera& c, uint32_t x, uint32_t y, float j1, float j2) {
  float fx = (((float)x + j1) * c.w_inv - 0.5f) * c.ar;
  float fy = 0.5f - ((float)y + j2) * c.h_inv;
  F3 p = normalize(f3(fx, fy, 0.8f));
  F3 rx = f3(p.x, c.cx * p.y - c.sx * p.z, c.sx * p.y + c.cx * p.z);        // rot_x, vec3.rs:108-119
  F3 ry = f3(c.cy * rx.x + c.sy * rx.z, rx.y, -c.sy * rx.x + c.cy * rx.z);  // rot_y, vec3.rs:95-106
  return make_ray(f3(c.ox, c.oy, c.oz), ry);
}

// ------------------------------------------------------------------ PNEE light choice
// PhotonTree::sample (photon_tree.rs:80-159) on the flattened octree.
struct Cell { float x0, y0, z0, x1, y1, z1; };
WPT_DEV uint32_t octree_child(Cell& b, F3 v) {   // photon_tree.rs:235-251
  float cx = 0.5f * (b.x0 + b.x1), cy = 0.5f * (b.y0 + b.y1), cz = 0.5f * (b.z0 + b.z1);
  uint32_t i = (v.x < cx ? 0u : 4u) + (v.y < cy ? 0u : 2u) + (v.z < cz ? 0u : 1u);
  if (v.x < cx) b.x1 = cx; else b.x0 = cx;
  if (v.y < cy) b.y1 = cy; else b.y0 = cy;
  if (v.z < cz) b.z1 = cz; else b.z0 = cz;
  return i;
}
WPT_DEV uint32_t tree_find_node(const DPhotonTree& t, uint32_t depth, F3 v) {   // find_node_cdf, photon_tree.rs:216-231
  Cell b = {-1024.0f, -1024.0f, -1024.0f, 1024.0f, 1024.0f, 1024.0f};
  uint32_t node = 0;
  for (;;) {
    uint32_t cb = __ldg(t.child_base + node);
    if (cb == 0xFFFFFFFFu || depth == 0) return node;
    node = cb + octree_child(b, v);
    depth--;
  }
}
WPT_DEV float tree_bin_prob(const DPhotonTree& t, uint32_t node, uint32_t i) {   // empirical_pdf.rs:64-75
  const float* cum = t.cum + (size_t)node * t.num_lights;
  if (i + 1 == t.num_lights) return 1.0f - __ldg(cum + i);
  return __ldg(cum + i + 1) - __ldg(cum + i);
}
WPT_DEV void axis_weight(float v, float c, float lo, float hi, float sz, float* w, float* w_adj, float* off) {   // photon_tree.rs:90-124
  if (v > c) { float lw = (hi - (v - sz * 0.5f)) / sz; *w = lw; *w_adj = 1.0f - lw; *off = 1.0f; }
  else { float rw = ((v + sz * 0.5f) - lo) / sz; *w = rw; *w_adj = 1.0f - rw; *off = -1.0f; }
}
__device__ __noinline__ void photon_sample(const DPhotonTree& t, Rng& rng, F3 v, uint32_t* light, float* pdf_out) {
  const float size = 1024.0f;
  if (v.x < -size || v.y < -size || v.z < -size || v.x > size || v.y > size || v.z > size) {
    *light = rng.range(0, t.num_lights);
    *pdf_out = 1.0f / (float)t.num_lights;
    return;
  }
  // find_leaf (photon_tree.rs:201-211)
  Cell b = {-size, -size, -size, size, size, size};
  uint32_t depth = 0, node = 0;
  for (;;) {
    uint32_t cb = __ldg(t.child_base + node);
    if (cb == 0xFFFFFFFFu) break;
    node = cb + octree_child(b, v);
    depth++;
  }
  float xs = b.x1 - b.x0, ys = b.y1 - b.y0, zs = b.z1 - b.z0;
  float wx, ax, ox, wy, ay, oy, wz, az, oz;
  axis_weight(v.x, 0.5f * (b.x0 + b.x1), b.x0, b.x1, xs, &wx, &ax, &ox);
  axis_weight(v.y, 0.5f * (b.y0 + b.y1), b.y0, b.y1, ys, &wy, &ay, &oy);
  axis_weight(v.z, 0.5f * (b.z0 + b.z1), b.z0, b.z1, zs, &wz, &az, &oz);
  bool self_x = rng.f32() <= wx;
  bool self_y = rng.f32() <= wy;
  bool self_z = rng.f32() <= wz;
  F3 sv = v;
  sv = sv + (self_x ? f3(0.0f, 0.0f, 0.0f) : ox * f3(xs, 0.0f, 0.0f));
  sv = sv + (self_y ? f3(0.0f, 0.0f, 0.0f) : oy * f3(0.0f, ys, 0.0f));
  sv = sv + (self_z ? f3(0.0f, 0.0f, 0.0f) : oz * f3(0.0f, 0.0f, zs));
  // EmpiricalPDF::sample (empirical_pdf.rs:43-61)
  uint32_t sn = tree_find_node(t, depth, sv);
  const float* cum = t.cum + (size_t)sn * t.num_lights;
  float r = rng.f32();
  uint32_t low = 0, high = t.num_lights;
  while (low + 1 < high) {
    uint32_t mid = (low + high) / 2;
    if (__ldg(cum + mid) <= r) low = mid; else high = mid;
  }
  uint32_t res = low;
  float ajx = xs * ox, ajy = ys * oy, ajz = zs * oz;
  float pdf = 0.0f;
  pdf += tree_bin_prob(t, tree_find_node(t, depth, v), res) * wx * wy * wz;
  pdf += tree_bin_prob(t, tree_find_node(t, depth, v + f3(ajx, 0.0f, 0.0f)), res) * ax * wy * wz;
  pdf += tree_bin_prob(t, tree_find_node(t, depth, v + f3(0.0f, ajy, 0.0f)), res) * wx * ay * wz;
  pdf += tree_bin_prob(t, tree_find_node(t, depth, v + f3(0.0f, 0.0f, ajz)), res) * wx * wy * az;
  pdf += tree_bin_prob(t, tree_find_node(t, depth, v + f3(ajx, ajy, 0.0f)), res) * ax * ay * wz;
  pdf += tree_bin_prob(t, tree_find_node(t, depth, v + f3(0.0f, ajy, ajz)), res) * wx * ay * az;
  pdf += tree_bin_prob(t, tree_find_node(t, depth, v + f3(ajx, 0.0f, ajz)), res) * ax * wy * az;
  pdf += tree_bin_prob(t, tree_find_node(t, depth, v + f3(ajx, ajy, ajz)), res) * ax * ay * az;
  *light = res;
  *pdf_out = pdf;
}

// Triangle::pick_random (triangle.rs:91-114) on light `li`
WPT_DEV void pick_random(const DScene& sc, uint32_t li, Rng& rng, F3* p, F3* n, F3* intensity, float* area, uint32_t* shape_id) {
  float4 na = __ldg(&sc.lights[li].n_area), in = __ldg(&sc.lights[li].intensity);
  uint32_t sid = __float_as_uint(in.w);
  const float4* q = reinterpret_cast<const float4*>(sc.shapes + sid);
  F3 v0 = xyz(__ldg(q)), v1 = xyz(__ldg(q + 1)), v2 = xyz(__ldg(q + 2));
  float r1 = rng.f32();
  float r2 = rng.f32();
  float r1s = sqrtf(r1);
  *p = (1.0f - r1s) * v0 + (r1s * (1.0f - r2)) * v1 + (r2 * r1s) * v2;
  F3 nn = xyz(na);
  if (rng.f32() > 0.5f) nn = -nn;
  *n = nn; *intensity = xyz(in); *area = na.w; *shape_id = sid;
}

// ------------------------------------------------------------------ shading of one hit
// One bounce of trace_original_color (tracer.rs:237-329) after Scene::trace returned shape
// `id` (-1: miss) for `ray`. Shared by the wavefront shade kernel and the persistent kernel so
// that both evaluate exactly the same f32 expressions in the same order.
struct PathRegs { F3 color, T; Rng rng; bool bounced; };
struct ShadeOut {
  bool finished;     // path ended at this vertex (miss / emitter): `color` is final
  bool survive;      // Russian roulette outcome (only meaningful if !finished)
  bool shadow;       // a shadow ray has to be traced; `contrib` is added if it is unoccluded
  F3 next_o, next_d; // the bounce ray
  F3 sh_o, sh_d; float sh_len; int sh_light; F3 contrib;
};
WPT_DEV void shade_hit(const RenderParams& rp, const Ray& ray, int id, PathRegs& ps, ShadeOut& out) {
  const bool has_nee = rp.render_type != 0;
  out.finished = false; out.survive = false; out.shadow = false;
  bool some = false; float t = 0.0f; F3 n = f3(0, 0, 0); uint32_t mat = 0;
  if (id >= 0) some = shape_trace_full(rp.scene.shapes, (uint32_t)id, ray, &t, &n, &mat);   // scene.rs:140
  if (!some) {   // tracer.rs:325-328
    ps.color = ps.color + ps.T * f3(rp.scene.bg_r, rp.scene.bg_g, rp.scene.bg_b);
    out.finished = true;
    return;
  }
  float4 mc = __ldg(&rp.scene.mats[mat].c);
  F3 hit_point = ray.o + t * ray.d;
  if (mc.w != 0.0f) {   // emissive, tracer.rs:245-254
    if (rp.light_debug ? !ps.bounced : (!has_nee || !ps.bounced)) ps.color = ps.color + ps.T * xyz(mc);
    out.finished = true;
    return;
  }
  // material.rs:97-118 cosine-weighted bounce
  float r1 = ps.rng.f32();
  float r2 = ps.rng.f32();
  float sa, ca;
  shared_sincos(2.0f * WPT_PI * r1, &sa, &ca);
  float x = ca * sqrtf(1.0f - r2);
  float y = sqrtf(r2);
  float z = sa * sqrtf(1.0f - r2);
  F3 xn = orthogonal(n);
  F3 zn = cross(n, xn);
  F3 wi = normalize(x * xn + y * n + z * zn);
  float pdf = dot(wi, n) / WPT_PI;
  const float inv_pi = 1.0f / WPT_PI;   // Color3 / f32 = self * (1/v), clamped (color3.rs:54-95)
  F3 brdf = f3(fminf(1.0f, fmaxf(0.0f, inv_pi * mc.x)), fminf(1.0f, fmaxf(0.0f, inv_pi * mc.y)), fminf(1.0f, fmaxf(0.0f, inv_pi * mc.z)));
  float cos_i = dot(wi, n);
  ps.T = ps.T * brdf * cos_i / pdf;   // tracer.rs:262
  out.next_o = hit_point + wi * WPT_EPSILON;
  out.next_d = wi;
  ps.bounced = true;
  if (has_nee) {   // tracer.rs:267-313
    uint32_t light_id; float chance;
    if (rp.render_type == 2) photon_sample(rp.photons, ps.rng, hit_point, &light_id, &chance);
    else { light_id = ps.rng.range(0, rp.scene.num_lights); chance = 1.0f / (float)rp.scene.num_lights; }
    F3 pl, ln, inten; float area; uint32_t lsid;
    pick_random(rp.scene, light_id, ps.rng, &pl, &ln, &inten, &area, &lsid);
    F3 to_light = pl - hit_point;
    float dsq = dot(to_light, to_light);
    float dlen = sqrtf(dsq);
    to_light = to_light / dlen;
    float cos_i2 = dot(to_light, n);
    float cos_o = dot(-to_light, ln);
    if (cos_i2 > 0.0f && cos_o > 0.0f) {
      if (rp.light_debug) ps.color = ps.color + ps.T * inten;
      else {
        float solid_angle = (area * cos_o) / dsq;
        out.contrib = ps.T * inten * solid_angle * cos_i2 * (1.0f / chance);
        out.sh_o = hit_point + to_light * WPT_EPSILON;   // scene.rs:108
        out.sh_d = to_light; out.sh_len = dlen; out.sh_light = (int)lsid;
        out.shadow = true;
      }
    }
  }
  // Russian roulette, tracer.rs:318-324
  float keep = fmaxf(fminf(fmaxf(fmaxf(ps.T.x, ps.T.y), ps.T.z), 0.9f), 0.1f);
  out.survive = ps.rng.f32() < keep;
  if (out.survive) ps.T = ps.T * (1.0f / keep);
}

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
  uint32_t row = i / w, col = i - row * w;
  pixel[i] = (y0 + row * world + rank) * W + x0 + col;
}
void launch_fill_pixels(uint32_t* pixel, uint32_t W, uint32_t x0, uint32_t y0, uint32_t w, uint32_t h, uint32_t rank, uint32_t world, cudaStream_t s) {
  uint32_t rows = h > rank ? (h - rank + world - 1) / world : 0;
  if (!rows || !w) return;
  k_fill_pixels<<<(w * rows + 255) / 256, 256, 0, s>>>(pixel, W, x0, y0, w, rows, rank, world);
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
        shade_hit(rp, ray, __float_as_int(h.y), ps, so);
        color = ps.color; T = ps.T; rng = ps.rng;
        if (so.finished) finished = true;
        else {
          ro.x = so.next_o.x; ro.y = so.next_o.y; ro.z = so.next_o.z;
          rd.x = so.next_d.x; rd.y = so.next_d.y; rd.z = so.next_d.z;
          flags |= SL_BOUNCED;
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
enum : int { PH_NEED = 0, PH_LOGIC = 1, PH_TRAV = 2, PH_DONE = 3 };
enum : int { ST_GEN = 0, ST_EXTEND = 1, ST_SHADOW = 2 };

template <int MINB>
__global__ void __launch_bounds__(MEGA_THREADS, MINB) k_mega(MegaParams P) {
  const DScene& sc = P.rp.scene;
  uint32_t stack_n[WPT_STACK]; float stack_d[WPT_STACK];
  const unsigned FULL = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u;
  int phase = PH_NEED, what = ST_GEN;
  uint32_t pix = 0, s = 0, s_end = 0;
  PathRegs ps; ps.color = f3(0, 0, 0); ps.T = f3(1, 1, 1); ps.rng.s = 1u; ps.bounced = false;
  Ray ray = make_ray(f3(0, 0, 0), f3(1, 1, 1));
  Trav tv; tv.lf = tv.cnt = 0; tv.sp = 0; tv.bound = 0; tv.best_id = -1; tv.inf_t = 0; tv.inf_id = -1; tv.visits = tv.prims = 0;
  F3 ext_o = f3(0, 0, 0), ext_d = f3(0, 0, 0), contrib = f3(0, 0, 0);
  float sh_len = 0.0f; int sh_light = -1; bool alive_after_shadow = false;
  uint32_t c_rays = 0, c_visits = 0, c_prims = 0, c_paths = 0;
  Accum acc{P.accum};
#ifdef MEGA_INSTR
  unsigned long long i_lp = 0, i_ll = 0, i_ts = 0, i_tl = 0, i_sh = 0;   // logic passes, logic lanes, trav steps, trav lanes, shade lanes
#endif

  for (;;) {
    // ---- pixel fetch: one atomic per warp
    unsigned need = __ballot_sync(FULL, phase == PH_NEED);
    if (need) {
      uint32_t base = 0;
      int leader = __ffs(need) - 1;
      if ((int)lane == leader) base = atomicAdd(P.work_counter, (uint32_t)__popc(need));
      base = __shfl_sync(FULL, base, leader);
      if (phase == PH_NEED) {
        uint32_t idx = base + (uint32_t)__popc(need & ((1u << lane) - 1u));
        if (idx < P.nslots) {
          pix = P.pixel[idx];
          uint32_t spp = P.spp_per_slot ? P.spp_per_slot[idx] : P.uniform_spp;
          s = __float_as_uint(P.accum[pix].w);   // samples accumulated so far = next sample index
          s_end = s + spp;
          what = ST_GEN; phase = PH_LOGIC;
        } else phase = PH_DONE;
      }
    }
    unsigned trav = __ballot_sync(FULL, phase == PH_TRAV);
    unsigned logic = __ballot_sync(FULL, phase == PH_LOGIC);
    if (!(trav | logic)) break;
    if (__popc(trav) >= (int)P.t_hi || !logic) {
      // ---- traversal burst, while-while: cheap inner steps until (almost) every lane of the
      // burst waits at a leaf, then one leaf step with all of them (triangle tests are the
      // expensive body: run them with as many lanes as possible)
      do {
        const int n_trav = __popc(trav);
        for (;;) {
          bool at_inner = phase == PH_TRAV && !trav_at_leaf(sc, tv);
          int n_inner = __popc(__ballot_sync(FULL, at_inner));
          if (n_inner == 0) break;
#ifdef MEGA_INSTR
          i_ts += 1; i_tl += n_inner;
#endif
          if (at_inner) {
            bool cont = sc.bvh_kind == 4 ? trav_inner4(sc, ray, tv, stack_n, stack_d) : trav_inner2(sc, ray, tv, stack_n, stack_d);
            if (!cont) phase = PH_LOGIC;
          }
          if (n_inner * P.t_inner <= n_trav) break;   // few lanes left at inner nodes: let them wait
        }
        if (phase == PH_TRAV && trav_at_leaf(sc, tv)) {
          bool cont = sc.bvh_kind == 4 ? trav_leaf4(sc, ray, tv, stack_n, stack_d) : trav_leaf2(sc, ray, tv, stack_n, stack_d);
          if (!cont) phase = PH_LOGIC;
        }
        trav = __ballot_sync(FULL, phase == PH_TRAV);
      } while (__popc(trav) >= (int)P.t_lo);
      continue;
    }
#ifdef MEGA_INSTR
    i_lp += 1; i_ll += __popc(logic); i_sh += __popc(__ballot_sync(FULL, phase == PH_LOGIC && what == ST_EXTEND));
#endif
    if (phase != PH_LOGIC) continue;
    // ---- logic pass
    bool start = false;
    if (what != ST_GEN) {
      GHit g = trav_result(tv);
      c_rays += 1; c_visits += g.visits; c_prims += g.prims;
      bool finish = false;
      if (what == ST_SHADOW) {   // Scene::shadow_ray, scene.rs:114-132
        bool occluded = g.id >= 0 && g.t < sh_len && g.id != sh_light;
        if (!occluded) ps.color = ps.color + contrib;
        if (alive_after_shadow) { ray = make_ray(ext_o, ext_d); what = ST_EXTEND; start = true; }
        else finish = true;
      } else {
        ShadeOut so;
        shade_hit(P.rp, ray, g.id, ps, so);
        if (so.finished) finish = true;
        else if (so.shadow) {
          ext_o = so.next_o; ext_d = so.next_d; contrib = so.contrib; sh_len = so.sh_len; sh_light = so.sh_light;
          alive_after_shadow = so.survive;
          ray = make_ray(so.sh_o, so.sh_d); what = ST_SHADOW; start = true;
        } else if (so.survive) { ray = make_ray(so.next_o, so.next_d); what = ST_EXTEND; start = true; }
        else finish = true;
      }
      if (finish) { acc.add(pix, ps.color); c_paths += 1; s += 1; what = ST_GEN; }
    }
    if (what == ST_GEN) {
      if (s < s_end) {   // tracer.rs:176-196 — sample s of this pixel on its own stream
        ps.rng.s = stream_seed(pix, s, STREAM_PATH, P.rp.base_seed);
        float j1 = ps.rng.f32();
        float j2 = ps.rng.f32();
        uint32_t py = pix / P.rp.W, px = pix - py * P.rp.W;
        ray = camera_ray(P.rp.cam, px, py, j1, j2);
        ps.color = f3(0, 0, 0); ps.T = f3(1.0f, 1.0f, 1.0f); ps.bounced = false;
        what = ST_EXTEND; start = true;
      } else phase = PH_NEED;
    }
    if (start && trav_begin(sc, ray, tv)) phase = PH_TRAV;
  }
  // ---- counters
  unsigned long long r = warp_sum_u64(c_rays), v = warp_sum_u64(c_visits), pr = warp_sum_u64(c_prims), pa = warp_sum_u64(c_paths);
  if (lane == 0 && r) {
    atomicAdd(&P.counters[0], r); atomicAdd(&P.counters[1], v); atomicAdd(&P.counters[2], pa); atomicAdd(&P.counters[3], pr);
  }
#ifdef MEGA_INSTR
  if (lane == 0) { atomicAdd(&P.counters[4], i_lp); atomicAdd(&P.counters[5], i_ll); atomicAdd(&P.counters[6], i_ts); atomicAdd(&P.counters[7], i_tl); }
  if (lane == 1) atomicAdd(&P.counters[8], i_sh);
#endif
}
void launch_mega(const MegaParams& P, int blocks_per_sm, cudaStream_t s) {
  if (!P.nslots) return;
  int grid = device_sm_count() * blocks_per_sm;
  int need = (int)((P.nslots + MEGA_THREADS - 1) / MEGA_THREADS);
  if (grid > need) grid = need;
  if (blocks_per_sm >= 6) k_mega<6><<<grid, MEGA_THREADS, 0, s>>>(P);
  else if (blocks_per_sm == 5) k_mega<5><<<grid, MEGA_THREADS, 0, s>>>(P);
  else if (blocks_per_sm == 3) k_mega<3><<<grid, MEGA_THREADS, 0, s>>>(P);
  else k_mega<4><<<grid, MEGA_THREADS, 0, s>>>(P);
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
    bool ok = g.id >= 0 && shape_trace_full(rp.scene.shapes, (uint32_t)g.id, ray, &t, &nn, &mat);
    normals[i * 3] = ok ? nn.x : 0.0f; normals[i * 3 + 1] = ok ? nn.y : 0.0f; normals[i * 3 + 2] = ok ? nn.z : 0.0f;
  }
}
void launch_trace_batch(const RenderParams& rp, const float* o, const float* d, uint64_t n, int32_t* ids, float* dist, uint32_t* visits, float* normals, cudaStream_t s) {
  if (!n) return;
  k_trace_batch<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(rp, o, d, n, ids, dist, visits, normals);
}

}  // namespace wpt
