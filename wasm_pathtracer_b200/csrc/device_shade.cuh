// Device-side shading shared by every engine: camera rays (tracer.rs:156-201), the PNEE light choice
// (photon_tree.rs:80-159), Triangle::pick_random and one bounce of trace_original_color (tracer.rs:237-329).
// All engines call the same shade_hit, so they evaluate the same f32 expressions in the same order.
#pragma once
#include "kernels.h"
#include "device_core.cuh"

namespace wpt {

// ------------------------------------------------------------------ helpers
WPT_DEV unsigned long long warp_sum_u64(unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
  return v;
}

// tracer.rs:176-191 — camera ray through pixel (x,y) with jitter (j1,j2)
WPT_DEV F3 camera_dir(const DCamera& c, uint32_t x, uint32_t y, float j1, float j2) {
  float fx = (((float)x + j1) * c.w_inv - 0.5f) * c.ar;
  float fy = 0.5f - ((float)y + j2) * c.h_inv;
  F3 p = normalize(f3(fx, fy, 0.8f));
  F3 rx = f3(p.x, c.cx * p.y - c.sx * p.z, c.sx * p.y + c.cx * p.z);        // rot_x, vec3.rs:108-119
  return f3(c.cy * rx.x + c.sy * rx.z, rx.y, -c.sy * rx.x + c.cy * rx.z);   // rot_y, vec3.rs:95-106
}
WPT_DEV Ray camera_ray(const DCamera& c, uint32_t x, uint32_t y, float j1, float j2) { return make_ray(f3(c.ox, c.oy, c.oz), camera_dir(c, x, y, j1, j2)); }

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
  // both branches of the reference end in one division by the cell size: select the numerator, divide once
  const bool up = v > c;
  const float num = up ? (hi - (v - sz * 0.5f)) : ((v + sz * 0.5f) - lo);
  const float ww = num / sz;
  *w = ww; *w_adj = 1.0f - ww; *off = up ? 1.0f : -1.0f;
}
// Same result as the reference's ten root-to-cell walks (find_leaf + find_node_cdf for the
// sampled cell + 8 for the interpolated pdf) with one: find_leaf. The eight query points are
// the corners of a box one cell wide, so each of the seven others lies in the face / edge /
// corner neighbour of the leaf: `nbr[leaf][27]` holds, per direction, the node the reference's
// walk to the leaf's depth ends in (same depth, or the shallower leaf covering it), computed on
// the host with the same f32 halving. A corner is only taken from the table if its coordinate
// really lies inside the neighbour interval (f32 rounding of v + size can put it on the far
// boundary; cells touching the +-1024 cube are excluded too) — otherwise the walk is redone.
WPT_DEV uint32_t tree_walk(const DPhotonTree& t, uint32_t depth, F3 q) {   // find_node_cdf, photon_tree.rs:216-231
  Cell c = {-1024.0f, -1024.0f, -1024.0f, 1024.0f, 1024.0f, 1024.0f};
  uint32_t nd = 0;
  for (;;) {
    uint32_t cb = __ldg(t.child_base + nd);
    if (cb == 0xFFFFFFFFu || depth == 0) return nd;
    nd = cb + octree_child(c, q);
    depth--;
  }
}
WPT_DEV void photon_sample_inl(const DPhotonTree& t, Rng& rng, F3 v, uint32_t* light, float* pdf_out) {
  const float size = 1024.0f;
  if (v.x < -size || v.y < -size || v.z < -size || v.x > size || v.y > size || v.z > size) {
    *light = rng.range(0, t.num_lights);
    *pdf_out = 1.0f / (float)t.num_lights;
    return;
  }
  // find_leaf (photon_tree.rs:201-211)
  // (tried and removed: a 64^3 entry table for the first six halvings of the +-1024 cube — 1.6 % at best, and its code cost the
  //  BVH4 + PNEE variant more than that, gpurun_out/r2j_entry.log, r2k_ab.log)
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
  // neighbour coordinates: v + (ajx, 0, 0) etc. (photon_tree.rs:141-156); adding +-0.0 keeps v
  const float X1 = v.x + xs * ox, Y1 = v.y + ys * oy, Z1 = v.z + zs * oz;
  // is the shifted coordinate strictly inside the adjacent interval (and that inside the cube)?
  bool okx = ox > 0.0f ? (X1 >= b.x1 && X1 < b.x1 + xs && b.x1 + xs <= size) : (X1 >= b.x0 - xs && X1 < b.x0 && b.x0 - xs >= -size);
  bool oky = oy > 0.0f ? (Y1 >= b.y1 && Y1 < b.y1 + ys && b.y1 + ys <= size) : (Y1 >= b.y0 - ys && Y1 < b.y0 && b.y0 - ys >= -size);
  bool okz = oz > 0.0f ? (Z1 >= b.z1 && Z1 < b.z1 + zs && b.z1 + zs <= size) : (Z1 >= b.z0 - zs && Z1 < b.z0 && b.z0 - zs >= -size);
  const int dx = ox > 0.0f ? 2 : 0, dy = oy > 0.0f ? 2 : 0, dz = oz > 0.0f ? 2 : 0;   // direction index 0,1,2 = -1,0,+1
  const uint32_t* nb = t.nbr + (size_t)node * 27;
  uint32_t n8[8];
  n8[0] = node;
#pragma unroll
  for (int k = 1; k < 8; k++) {
    const bool bx = k & 1, by = k & 2, bz = k & 4;
    bool ok = (!bx || okx) && (!by || oky) && (!bz || okz);
    if (ok) n8[k] = __ldg(nb + (bx ? dx : 1) + 3 * (by ? dy : 1) + 9 * (bz ? dz : 1));
    else n8[k] = tree_walk(t, depth, f3(bx ? X1 : v.x, by ? Y1 : v.y, bz ? Z1 : v.z));
  }
  // EmpiricalPDF::sample (empirical_pdf.rs:43-61) on the sampled cell
#pragma unroll
  for (int k = 0; k < 8; k++) WPT_CHECK(n8[k] < t.num_nodes);
  uint32_t sel = (self_x ? 0u : 1u) + (self_y ? 0u : 2u) + (self_z ? 0u : 4u);
  uint32_t sn = n8[0];
#pragma unroll
  for (int k = 1; k < 8; k++) if (sel == (uint32_t)k) sn = n8[k];
  const float* cum = t.cum + (size_t)sn * t.num_lights;
  float r = rng.f32();
  uint32_t low = 0, high = t.num_lights;
  while (low + 1 < high) {
    uint32_t mid = (low + high) / 2;
    if (__ldg(cum + mid) <= r) low = mid; else high = mid;
  }
  uint32_t res = low;
  // EmpiricalPDF::bin_prob (empirical_pdf.rs:64-75) of bin `res` in each of the eight cells: cum[res + 1] - cum[res], with
  // 1.0 in place of cum[res + 1] for the last bin — the same subtraction either way, so the test is hoisted out
  const bool last = res + 1 == t.num_lights;
  const size_t L = t.num_lights;
  // (tried: one 8-byte load per cell when there are two lights, as in the bunny scene — BVH2 + PNEE 24.1 -> 25.2 ms, BVH4 + PNEE
  //  30.8 -> 32.9 ms: slower, gpurun_out/r2c_ab.log)
  auto bp = [&](uint32_t nd) { const float* c = t.cum + (size_t)nd * L + res; float a = __ldg(c); float hi = last ? 1.0f : __ldg(c + 1); return hi - a; };
  float pdf = 0.0f;   // the reference's order of the eight terms (photon_tree.rs:149-156)
  pdf += bp(n8[0]) * wx * wy * wz;
  pdf += bp(n8[1]) * ax * wy * wz;
  pdf += bp(n8[2]) * wx * ay * wz;
  pdf += bp(n8[4]) * wx * wy * az;
  pdf += bp(n8[3]) * ax * ay * wz;
  pdf += bp(n8[6]) * wx * ay * az;
  pdf += bp(n8[5]) * ax * wy * az;
  pdf += bp(n8[7]) * ax * ay * az;
  *light = res;
  *pdf_out = pdf;
}
// out-of-line copy for the kernels that take the render type at run time (k_shade, k_pool, the sample-batch probe)
static __device__ __noinline__ void photon_sample(const DPhotonTree& t, Rng& rng, F3 v, uint32_t* light, float* pdf_out) { photon_sample_inl(t, rng, v, light, pdf_out); }

// Triangle::pick_random (triangle.rs:91-114) on light `li`
WPT_DEV void pick_random(const DScene& sc, uint32_t li, Rng& rng, F3* p, F3* n, F3* intensity, float* area, uint32_t* shape_id) {
  WPT_CHECK(li < sc.num_lights);
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
// RT: the render type when it is known at compile time (0 NoNEE, 1 NormalNEE, 2 PNEE: k_mega is instantiated per type — no
// photon code in the NEE kernels, photon_sample inlined in the PNEE kernel: 9 % / 15 % faster than one kernel that branches
// and calls), 3 = decided at run time (k_shade, k_pool: out-of-line photon_sample)
template <int KIND, int RT = 3, bool PRE = false>
WPT_DEV void shade_hit(const RenderParams& rp, const Ray& ray, int id, float t_hit, PathRegs& ps, ShadeOut& out, const TorusPre* pre = nullptr) {
  const bool has_nee = RT == 3 ? rp.render_type != 0 : RT != 0;
  out.finished = false; out.survive = false; out.shadow = false;
  bool some = false; float t = 0.0f; F3 n = f3(0, 0, 0); uint32_t mat = 0;
  bool entering = true; float2 uv = make_float2(0.0f, 0.0f);
  WPT_CHECK(id < (int)rp.scene.num_shapes);
  if (id >= 0) {   // scene.rs:140
    if (KIND == K_SIMPLE) { shape_hit_normal_tri_plane(rp.scene.shapes, (uint32_t)id, ray, &n, &mat); t = t_hit; some = true; }
    else some = shape_trace_full<KIND, PRE>(rp.scene.shapes, (uint32_t)id, ray, &t, &n, &mat, &entering, &uv, t_hit, pre);
  }
  if (!some) {   // tracer.rs:325-328
    ps.color = ps.color + ps.T * f3(rp.scene.bg_r, rp.scene.bg_g, rp.scene.bg_b);
    out.finished = true;
    return;
  }
  float4 mc = __ldg(&rp.scene.mats[mat].c);
  F3 hit_point = ray.o + t * ray.d;
  if (mc.w == (float)MAT_EMISSIVE) {   // emissive, tracer.rs:245-254
    if (rp.light_debug ? !ps.bounced : (!has_nee || !ps.bounced)) ps.color = ps.color + ps.T * xyz(mc);
    out.finished = true;
    return;
  }
  if (KIND == K_EXT && mc.w != (float)MAT_DIFFUSE) {   // ---- extension materials (parity unpinned)
    float4 mp = __ldg(&rp.scene.mats[mat].p);
    if (mc.w == (float)MAT_DIFFUSE_TEX) {   // Texture::at, texture.rs:23-31 (`as u32` saturates)
      const DTexture tx = rp.scene.tex[__float_as_uint(mp.y)];
      float fu = floorf(uv.x * (float)tx.width), fv = floorf(uv.y * (float)tx.height);
      uint32_t iu = !(fu > 0.0f) ? 0u : (fu >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)fu);
      uint32_t iv = !(fv > 0.0f) ? 0u : (fv >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)fv);
      const uint8_t* px = tx.rgb + (size_t)((iv % tx.height) * tx.width + (iu % tx.width)) * 3;
      mc.x = (float)px[0] / 255.0f; mc.y = (float)px[1] / 255.0f; mc.z = (float)px[2] / 255.0f;
    } else if (mc.w == (float)MAT_REFRACT || ps.rng.f32() < mp.x) {   // specular bounce (Reflect draws its share first)
      F3 d = ray.d, wi;
      float cos_in = dot(-d, n);
      if (mc.w == (float)MAT_REFLECT) {
        wi = (2.0f * cos_in) * n - (-d);   // Vec3::reflect, vec3.rs:85-87
        ps.T = ps.T * xyz(mc);
      } else {
        float n1 = entering ? 1.0f : mp.x, n2 = entering ? mp.x : 1.0f;
        if (!entering) ps.T = ps.T * f3(shared_exp_neg(mc.x * t), shared_exp_neg(mc.y * t), shared_exp_neg(mc.z * t));   // Beer's law
        float r0 = (n1 - n2) / (n1 + n2); r0 = r0 * r0;   // Schlick with total internal reflection
        float cosx = cos_in; bool tir = false;
        if (n1 > n2) { float nr = n1 / n2; float sin2 = nr * nr * (1.0f - cosx * cosx); if (sin2 > 1.0f) tir = true; else cosx = sqrtf(1.0f - sin2); }
        float x = 1.0f - cosx;
        float fres = tir ? 1.0f : r0 + (1.0f - r0) * x * x * x * x * x;
        if (ps.rng.f32() < fres) wi = (2.0f * cos_in) * n - (-d);
        else {
          float eta = n1 / n2;
          float k = 1.0f - eta * eta * (1.0f - cos_in * cos_in);
          wi = normalize(eta * d + (eta * cos_in - sqrtf(fmaxf(k, 0.0f))) * n);
        }
      }
      out.next_o = hit_point + wi * WPT_EPSILON;
      out.next_d = wi;
      ps.bounced = false;   // a light seen through a specular bounce is not covered by NEE
      float keep = fmaxf(fminf(fmaxf(fmaxf(ps.T.x, ps.T.y), ps.T.z), 0.9f), 0.1f);
      out.survive = ps.rng.f32() < keep;
      if (out.survive) ps.T = ps.T * (1.0f / keep);
      return;
    }
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
    if (RT == 2) photon_sample_inl(rp.photons, ps.rng, hit_point, &light_id, &chance);
    else if (RT == 3 && rp.render_type == 2) photon_sample(rp.photons, ps.rng, hit_point, &light_id, &chance);
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

}  // namespace wpt
