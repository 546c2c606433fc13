// Host-side scene setup, BVH2 build, BVH4 collapse, flattening. See host_scene.h.
// Compiled with -ffp-contract=off: all f32 arithmetic here must round exactly like the
// reference's Rust (no FMA contraction), because node boxes, split decisions and the
// precomputed triangle normals feed bit-exact comparisons on the device.
#include "host_scene.h"
#include <vector_functions.h>
#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace wpt {

static inline float mn(float a, float b) { return std::fmin(a, b); }
static inline float mx(float a, float b) { return std::fmax(a, b); }

// ------------------------------------------------------------------ tiny f32 vector helpers
static inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 crossv(V3 a, V3 t) { return V3{a.y * t.z - a.z * t.y, a.z * t.x - a.x * t.z, a.x * t.y - a.y * t.x}; }
static inline float dotv(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float lenv(V3 a) { return std::sqrt(dotv(a, a)); }
static inline V3 normv(V3 a) { float s = 1.0f / lenv(a); return V3{s * a.x, s * a.y, s * a.z}; }   // vec3.rs:27-29

// xorshift32 (rng.rs:40-47) — only used for the museum colour shuffle (scenes.rs:30-40)
struct HostRng {
  uint32_t s = 0xBABABEBEu;
  uint32_t u32() { uint32_t x = s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; s = x; return x; }
  float f32() { return (float)u32() * (1.0f / 4294967296.0f); }
  size_t range(size_t lo, size_t hi) {   // rng.rs:25-38
    if (hi == lo + 1) return 0;
    float f = f32();
    if (f == 1.0f) return hi - 1;
    return (size_t)std::floor(f * (float)(hi - lo)) + lo;
  }
};

static HostShape tri(V3 a, V3 b, V3 c, uint32_t mat) {
  HostShape s{}; s.type = SH_TRIANGLE; s.mat = mat;
  float v[9] = {a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z};
  std::memcpy(s.p, v, sizeof v);
  return s;
}
static HostShape plane(V3 loc, V3 n, uint32_t mat) {
  HostShape s{}; s.type = SH_PLANE; s.mat = mat;
  float v[9] = {loc.x, loc.y, loc.z, n.x, n.y, n.z, 0, 0, 0};
  std::memcpy(s.p, v, sizeof v);
  return s;
}
static HostShape torus(V3 loc, float R, float r, uint32_t mat) {
  HostShape s{}; s.type = SH_TORUS; s.mat = mat;
  float v[9] = {loc.x, loc.y, loc.z, R, r, 0, 0, 0, 0};
  std::memcpy(s.p, v, sizeof v);
  return s;
}
static HostShape aarect(float x0, float x1, float y0, float y1, float z0, float z1, uint32_t mat) {
  HostShape s{}; s.type = SH_AARECT; s.mat = mat;
  float v[9] = {x0, y0, z0, x1, y1, z1, 0, 0, 0};
  std::memcpy(s.p, v, sizeof v);
  return s;
}
static HostShape sphere(V3 loc, float radius, uint32_t mat) {   // sphere.rs:17-22
  HostShape s{}; s.type = SH_SPHERE; s.mat = mat;
  float v[9] = {loc.x, loc.y, loc.z, radius, 0, 0, 0, 0, 0};
  std::memcpy(s.p, v, sizeof v);
  return s;
}
static HostShape square(V3 loc, float size, uint32_t mat) {     // square.rs:17-22
  HostShape s{}; s.type = SH_SQUARE; s.mat = mat;
  float v[9] = {loc.x, loc.y, loc.z, size, 0, 0, 0, 0, 0};
  std::memcpy(s.p, v, sizeof v);
  return s;
}
static inline float clamp01(float v) { return mn(1.0f, mx(0.0f, v)); }   // Color3::new, color3.rs:32-38
static uint32_t add_mat(std::vector<HostMaterial>& mats, float r, float g, float b, bool emissive) {
  if (!emissive) { r = clamp01(r); g = clamp01(g); b = clamp01(b); }
  HostMaterial m{}; m.r = r; m.g = g; m.b = b; m.emissive = emissive; m.kind = emissive ? MAT_EMISSIVE : MAT_DIFFUSE;
  mats.push_back(m);
  return (uint32_t)mats.size() - 1;
}
static uint32_t add_mat_ext(std::vector<HostMaterial>& mats, uint32_t kind, float r, float g, float b, float param, uint32_t tex) {
  HostMaterial m{}; m.r = r; m.g = g; m.b = b; m.emissive = false; m.kind = kind; m.param = param; m.tex = tex;
  if (kind == MAT_REFLECT || kind == MAT_DIFFUSE_TEX) { m.r = clamp01(r); m.g = clamp01(g); m.b = clamp01(b); }
  mats.push_back(m);
  return (uint32_t)mats.size() - 1;
}

// scenes.rs:54-68
static void museum_lights(std::vector<HostShape>& dst, float x, float z, uint32_t mat) {
  for (int side = 0; side < 2; side++) {
    // (z + 2.8, z + 2.5) on the far side, (z - 2.8, z - 2.5) on the near side
    float za = side == 0 ? z + 2.8f : z - 2.8f, zb = side == 0 ? z + 2.5f : z - 2.5f;
    V3 lc1{x - 1.0f, 0.0f, za}, lc2{x + 1.0f, 0.0f, za}, lc3{x + 1.0f, 1.0f, zb}, lc4{x - 1.0f, 1.0f, zb};
    dst.push_back(tri(lc3, lc2, lc1, mat));
    dst.push_back(tri(lc4, lc3, lc1, mat));
  }
}

// scenes.rs:15-52
void scene_museum(std::vector<HostShape>& shapes, std::vector<HostMaterial>& mats) {
  shapes.clear(); mats.clear();
  uint32_t grey = add_mat(mats, 0.7f, 0.7f, 0.7f, false);
  uint32_t white = add_mat(mats, 1.0f, 1.0f, 1.0f, false);
  shapes.push_back(plane(V3{0.0f, -1.0f, 0.0f}, V3{0.0f, 1.0f, 0.0f}, grey));
  const float xs[9] = {-16.0f, -12.0f, -8.0f, -4.0f, 0.0f, 4.0f, 8.0f, 12.0f, 16.0f};
  float colors[9][3] = {{1.0f, 0.3f, 0.3f}, {0.0f, 1.0f, 1.0f}, {0.3f, 0.3f, 1.0f}, {1.0f, 0.0f, 0.0f}, {0.0f, 1.0f, 0.0f},
                        {0.0f, 0.0f, 1.0f}, {1.0f, 0.0f, 1.0f}, {1.0f, 1.0f, 0.0f}, {0.3f, 1.0f, 0.3f}};
  HostRng rng;
  rng.f32();
  rng.f32();
  const float rows[3] = {-7.5f, 0.0f, 7.5f};
  for (int r = 0; r < 3; r++) {
    for (int i = 0; i < 9; i++) {
      shapes.push_back(torus(V3{xs[i], -0.5f, rows[r]}, 1.3f, 0.3f, white));
      // colour.to_vec3() * 2.5 (scenes.rs:37): Vec3 * f32 = multiplier * component
      uint32_t lm = add_mat(mats, 2.5f * colors[i][0], 2.5f * colors[i][1], 2.5f * colors[i][2], true);
      museum_lights(shapes, xs[i], rows[r], lm);
    }
    for (size_t i = 0; i < 9; i++) {   // Rng::shuffle, rng.rs:70-75
      size_t j = rng.range(0, 9);
      for (int c = 0; c < 3; c++) std::swap(colors[i][c], colors[j][c]);
    }
  }
  const float ws[8] = {-14.0f, -10.0f, -6.0f, -2.0f, 2.0f, 6.0f, 10.0f, 14.0f};
  for (float x : ws) shapes.push_back(aarect(x - 0.1f, x + 0.1f, -1.0f, 2.0f, -20.0f, 20.0f, grey));
  shapes.push_back(aarect(-20.0f, 20.0f, -1.0f, 2.0f, 3.75f - 0.1f, 3.75f + 0.1f, grey));
  shapes.push_back(aarect(-20.0f, 20.0f, -1.0f, 2.0f, -3.75f - 0.1f, -3.75f + 0.1f, grey));
}

// scenes.rs:75-111. Material slots: 0 floor, 1 wall, 2 mesh, 3 light
void scene_bunny(const std::vector<HostShape>* mesh, std::vector<HostShape>& shapes, std::vector<HostMaterial>& mats) {
  shapes.clear(); mats.clear();
  uint32_t floor_m = add_mat(mats, 1.0f, 1.0f, 1.0f, false);
  uint32_t wall_m = add_mat(mats, 0.8f, 1.0f, 0.8f, false);
  add_mat(mats, 1.0f, 0.4f, 0.4f, false);   // mesh material, wasm_interface.rs:300
  uint32_t light_m = add_mat(mats, 16.0f, 16.0f, 16.0f, true);
  shapes.push_back(plane(V3{0.0f, -1.0f, 0.0f}, V3{0.0f, 1.0f, 0.0f}, floor_m));
  shapes.push_back(plane(V3{0.0f, 0.0f, 13.0f}, V3{0.0f, 0.0f, -1.0f}, wall_m));
  if (mesh) shapes.insert(shapes.end(), mesh->begin(), mesh->end());
  V3 lc1{-1.0f, 7.0f, 0.0f}, lc2{1.0f, 7.0f, 0.0f}, lc3{1.0f, 7.0f, 2.0f}, lc4{-1.0f, 7.0f, 2.0f};
  shapes.push_back(tri(lc3, lc2, lc1, light_m));
  shapes.push_back(tri(lc4, lc3, lc1, light_m));
}

// Extension scene 256 — scenes.rs:113-130 with an emissive quad instead of the removed point light
void scene_whitted(bool with_floor, std::vector<HostShape>& shapes, std::vector<HostMaterial>& mats, float bg[3]) {
  shapes.clear(); mats.clear();
  bg[0] = 135.0f / 255.0f; bg[1] = 206.0f / 255.0f; bg[2] = 250.0f / 255.0f;
  if (with_floor) shapes.push_back(square(V3{0.0f, -1.0f, 4.0f}, 8.0f, add_mat_ext(mats, MAT_DIFFUSE_TEX, 1.0f, 1.0f, 1.0f, 0.0f, 0)));
  shapes.push_back(sphere(V3{-1.3f, 1.0f, -0.2f}, 0.7f, add_mat_ext(mats, MAT_REFRACT, 0.5f, 1.0f, 0.5f, 1.02f, 0)));
  shapes.push_back(sphere(V3{-0.4f, 0.0f, 1.0f}, 0.6f, add_mat_ext(mats, MAT_REFLECT, 1.0f, 1.0f, 1.0f, 0.3f, 0)));
  uint32_t light_m = add_mat(mats, 16.0f, 16.0f, 16.0f, true);
  V3 lc1{-1.0f, 6.0f, -3.0f}, lc2{1.0f, 6.0f, -3.0f}, lc3{1.0f, 6.0f, -1.0f}, lc4{-1.0f, 6.0f, -1.0f};
  shapes.push_back(tri(lc3, lc2, lc1, light_m));
  shapes.push_back(tri(lc4, lc3, lc1, light_m));
}

// wasm_interface.rs:297-313: v * 0.5, then translate by (0,0,5)
std::vector<HostShape> mesh_triangles(const float* verts, size_t num_vertices, uint32_t mat) {
  size_t nt = num_vertices / 3;
  std::vector<HostShape> out;
  out.reserve(nt);
  for (size_t i = 0; i < nt; i++) {
    const float* p = verts + i * 9;
    V3 v[3];
    for (int k = 0; k < 3; k++) v[k] = V3{0.5f * p[k * 3] + 0.0f, 0.5f * p[k * 3 + 1] + 0.0f, 0.5f * p[k * 3 + 2] + 5.0f};
    out.push_back(tri(v[0], v[1], v[2], mat));
  }
  return out;
}

// ------------------------------------------------------------------ bounds / centroids
static bool shape_box(const HostShape& s, Box* b) {
  switch (s.type) {
    case SH_TRIANGLE: {   // triangle.rs:48-66
      const float pad = 0.1f * WPT_EPSILON;
      for (int c = 0; c < 3; c++) {
        b->lo[c] = mn(mn(s.p[c], s.p[3 + c]), s.p[6 + c]) - pad;
        b->hi[c] = mx(mx(s.p[c], s.p[3 + c]), s.p[6 + c]) + pad;
      }
      return true;
    }
    case SH_TORUS: {      // torus.rs:33-52
      float r = s.p[3] + s.p[4];
      b->lo[0] = s.p[0] - r; b->hi[0] = s.p[0] + r;
      b->lo[1] = s.p[1] - s.p[4]; b->hi[1] = s.p[1] + s.p[4];
      b->lo[2] = s.p[2] - r; b->hi[2] = s.p[2] + r;
      return true;
    }
    case SH_AARECT:       // aa_rect.rs:57-67
      for (int c = 0; c < 3; c++) { b->lo[c] = s.p[c]; b->hi[c] = s.p[3 + c]; }
      return true;
    case SH_SPHERE:       // sphere.rs:31-36
      for (int c = 0; c < 3; c++) { b->lo[c] = s.p[c] - s.p[3]; b->hi[c] = s.p[c] + s.p[3]; }
      return true;
    case SH_SQUARE: {     // square.rs:31-44
      float hs = s.p[3] * 0.5f;
      b->lo[0] = s.p[0] - hs; b->hi[0] = s.p[0] + hs;
      b->lo[1] = s.p[1]; b->hi[1] = s.p[1];
      b->lo[2] = s.p[2] - hs; b->hi[2] = s.p[2] + hs;
      return true;
    }
    default: return false;
  }
}
static V3 shape_centroid(const HostShape& s, const Box& b) {
  if (s.type == SH_TORUS || s.type == SH_SPHERE || s.type == SH_SQUARE) return V3{s.p[0], s.p[1], s.p[2]};   // torus.rs:28-30, sphere.rs:26-28, square.rs:26-28
  // triangle: AABB centre (ray.rs:77-83); aa_rect: aa_rect.rs:48-54 — the same expression
  return V3{0.5f * (b.lo[0] + b.hi[0]), 0.5f * (b.lo[1] + b.hi[1]), 0.5f * (b.lo[2] + b.hi[2])};
}
static inline Box join(const Box& a, const Box& b) {
  Box r;
  for (int c = 0; c < 3; c++) { r.lo[c] = mn(a.lo[c], b.lo[c]); r.hi[c] = mx(a.hi[c], b.hi[c]); }
  return r;
}
static inline float surface(const Box& b) {   // aabb.rs:72-78
  float xs = b.hi[0] - b.lo[0], ys = b.hi[1] - b.lo[1], zs = b.hi[2] - b.lo[2];
  return 2.0f * (xs * ys + xs * zs + ys * zs);
}

// ------------------------------------------------------------------ BVH2 (bvh.rs)
namespace {
struct Builder {
  const std::vector<Box>& boxes;
  const std::vector<V3>& cents;
  std::vector<uint32_t> order, scratch;
  std::vector<uint8_t> bin_of;
  std::vector<HostBVH2Node>& out;
  uint32_t num_bins;

  Builder(const std::vector<Box>& b, const std::vector<V3>& c, std::vector<HostBVH2Node>& o, uint32_t nb)
      : boxes(b), cents(c), order(b.size()), scratch(b.size()), bin_of(b.size()), out(o), num_bins(nb) {
    for (size_t i = 0; i < order.size(); i++) order[i] = (uint32_t)i;
  }
  Box range_box(size_t off, size_t len) const {
    Box r = boxes[order[off]];
    for (size_t i = 1; i < len; i++) r = join(r, boxes[order[off + i]]);
    return r;
  }
  static float axis_of(const V3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

  HostBVH2Node subdivide(size_t off, size_t len, const Box& parent) {
    HostBVH2Node leaf{};
    leaf.left_first = (uint32_t)off; leaf.count = (uint32_t)len;
    if (len <= 1) { leaf.box = range_box(off, len); return leaf; }
    // bvh.rs:282-303: longest axis of the box handed down, ties prefer z then y
    float xs = parent.hi[0] - parent.lo[0], ys = parent.hi[1] - parent.lo[1], zs = parent.hi[2] - parent.lo[2];
    int axis = (xs > ys) ? ((xs > zs) ? 0 : 2) : ((ys > zs) ? 1 : 2);
    // bvh.rs:412-437: uniform bins over the centroid extent
    float vmin = axis_of(cents[order[off]], axis), vmax = vmin;
    for (size_t i = 1; i < len; i++) { float v = axis_of(cents[order[off + i]], axis); vmin = mn(vmin, v); vmax = mx(vmax, v); }
    if (vmin == vmax) { leaf.box = range_box(off, len); return leaf; }
    const uint32_t NB = num_bins;
    std::vector<Box> bbox(NB);
    std::vector<uint32_t> bcnt(NB, 0);
    float seg = (vmax - vmin) / (float)NB;
    for (size_t i = 0; i < len; i++) {
      uint32_t id = order[off + i];
      float q = std::floor((axis_of(cents[id], axis) - vmin) / seg);
      uint32_t b = q >= 0.0f ? (q >= (float)NB ? NB - 1 : (uint32_t)q) : 0;
      if (b > NB - 1) b = NB - 1;
      bin_of[off + i] = (uint8_t)b;
      bbox[b] = bcnt[b] ? join(bbox[b], boxes[id]) : boxes[id];
      bcnt[b]++;
    }
    // bvh.rs:328-369: greedy two-ended sweep
    uint32_t l = 0, r = NB - 1;
    Box lb = bbox[l], rb = bbox[r];
    uint32_t lc = bcnt[l], rc = bcnt[r];
    Box lnb = bcnt[l + 1] ? join(lb, bbox[l + 1]) : lb;
    Box rnb = bcnt[r - 1] ? join(rb, bbox[r - 1]) : rb;
    uint32_t lnc = lc + bcnt[l + 1], rnc = rc + bcnt[r - 1];
    while (l + 1 < r) {
      if ((surface(lnb) * (float)lnc + surface(rb) * (float)rc) < (surface(lb) * (float)lc + surface(rnb) * (float)rnc)) {
        l++; lb = lnb; lc = lnc;
        if (l + 1 < r) { lnb = bcnt[l + 1] ? join(lb, bbox[l + 1]) : lb; lnc = lc + bcnt[l + 1]; }
      } else {
        r--; rb = rnb; rc = rnc;
        if (l + 1 < r) { rnb = bcnt[r - 1] ? join(rb, bbox[r - 1]) : rb; rnc = rc + bcnt[r - 1]; }
      }
    }
    // bvh.rs:264-273: accept only if it beats the unsplit cost
    float utility = surface(lb) * (float)lc + surface(rb) * (float)(len - lc);
    Box joined = join(lb, rb);
    if (!(utility < surface(joined) * (float)len)) { leaf.box = joined; return leaf; }
    // bvh.rs:462-470: write the bins back in bin order, stable inside a bin = counting sort
    std::vector<uint32_t> start(NB + 1, 0);
    for (uint32_t b = 0; b < NB; b++) start[b + 1] = start[b] + bcnt[b];
    for (size_t i = 0; i < len; i++) scratch[off + start[bin_of[off + i]]++] = order[off + i];
    std::copy(scratch.begin() + off, scratch.begin() + off + len, order.begin() + off);
    // bvh.rs:225-230: reserve the sibling pair, then left subtree fully before right
    size_t pair = out.size();
    out.push_back(HostBVH2Node{});
    out.push_back(HostBVH2Node{});
    HostBVH2Node ln = subdivide(off, lc, lb);
    out[pair] = ln;
    HostBVH2Node rn = subdivide(off + lc, len - lc, rb);
    out[pair + 1] = rn;
    HostBVH2Node inner{};
    inner.box = joined; inner.left_first = (uint32_t)pair; inner.count = 0;
    return inner;
  }
};

uint32_t depth2_of(const std::vector<HostBVH2Node>& n, uint32_t i) {
  if (n[i].count) return 0;
  return 1 + std::max(depth2_of(n, n[i].left_first), depth2_of(n, n[i].left_first + 1));
}

// ---------------------------------------------------------------- BVH4 (bvh4.rs)
// Tree-cut DP (bvh4.rs:244-281) evaluated bottom-up: children always have larger indices
// than their parent (bvh.rs:225-230), so one reverse sweep fills every cost vector. Costs are
// small integers (exact in the reference's f32).
struct Collapser {
  const std::vector<HostBVH2Node>& b2;
  std::vector<std::array<uint32_t, 4>> m;
  std::vector<HostBVH4Node>& dst;
  Collapser(const std::vector<HostBVH2Node>& b, std::vector<HostBVH4Node>& d) : b2(b), m(b.size()), dst(d) {
    for (size_t k = b2.size(); k-- > 0;) {
      if (k == 1 || b2[k].count) continue;   // pad slot / leaf
      uint32_t L = b2[k].left_first, R = L + 1;
      std::array<uint32_t, 4> c = {UINT32_MAX, UINT32_MAX, UINT32_MAX, UINT32_MAX};
      for (uint32_t t = 2; t <= 4; t++) {
        for (uint32_t i = 1; i < t; i++) c[t - 1] = std::min(c[t - 1], flat(L, i) + flat(R, t - i));
        c[0] = std::min(c[0], 1 + c[t - 1]);
      }
      m[k] = c;
    }
  }
  uint32_t flat(uint32_t node, uint32_t cut) const {   // bvh4.rs:228-240
    if (b2[node].count) return 1;
    uint32_t v = m[node][0];
    for (uint32_t i = 1; i < cut; i++) v = std::min(v, m[node][i]);
    return v;
  }
  uint32_t find_t(uint32_t node, uint32_t cut) const {   // bvh4.rs:189-205 (first minimum wins)
    if (b2[node].count) return 1;
    uint32_t t_min = 1, val = m[node][0];
    for (uint32_t t = 2; t <= cut; t++) if (m[node][t - 1] < val) { t_min = t; val = m[node][t - 1]; }
    return t_min;
  }
  uint32_t find_i(uint32_t L, uint32_t R, uint32_t t) const {   // bvh4.rs:210-224
    uint32_t i_min = 1, val = flat(L, 1) + flat(R, t - 1);
    for (uint32_t i = 2; i < t; i++) { uint32_t v = flat(L, i) + flat(R, t - i); if (v < val) { i_min = i; val = v; } }
    return i_min;
  }
  struct Cut { Box box[4]; int32_t id[4]; uint32_t n = 0; void push(const Box& b, int32_t i) { box[n] = b; id[n] = i; n++; } };
  void emit(uint32_t node, uint32_t cut, Cut& out) {   // bvh4.rs:127-185
    const HostBVH2Node& nd = b2[node];
    if (nd.count) {
      if (nd.count > 15) throw std::runtime_error("BVH4: leaf with more than 15 shapes cannot be encoded");
      out.push(nd.box, (int32_t)(0x80000000u | (nd.count << 27) | nd.left_first));
      return;
    }
    uint32_t L = nd.left_first, R = L + 1;
    uint32_t t = find_t(node, cut);
    if (t == 1) {
      size_t index = dst.size();
      dst.push_back(HostBVH4Node{});
      uint32_t i_min = find_i(L, R, 4);
      Cut kids;
      emit(L, i_min, kids);
      emit(R, 4 - i_min, kids);
      HostBVH4Node n4{};
      for (int k = 0; k < 4; k++) n4.children[k] = INT32_MIN;
      Box hull = kids.box[0];
      for (uint32_t k = 0; k < kids.n; k++) { n4.child[k] = kids.box[k]; n4.children[k] = kids.id[k]; if (k) hull = join(hull, kids.box[k]); }
      n4.num_children = kids.n;
      dst[index] = n4;
      out.push(hull, (int32_t)index);
      return;
    }
    uint32_t i_min = find_i(L, R, t);
    emit(L, i_min, out);
    emit(R, t - i_min, out);
  }
  void run() {   // bvh4.rs:37-70
    Cut res;
    emit(0, 4, res);
    if (res.n > 1) {
      dst.clear();
      dst.push_back(HostBVH4Node{});
      Cut res2;
      emit(0, 4, res2);
      HostBVH4Node root{};
      for (int k = 0; k < 4; k++) root.children[k] = 0;
      for (uint32_t k = 0; k < res2.n; k++) { root.child[k] = res2.box[k]; root.children[k] = res2.id[k]; }
      root.num_children = res2.n;
      dst[0] = root;
    } else if (res.id[0] != 0) throw std::runtime_error("BVH4: root is a leaf (bvh4.rs:67)");
  }
};
uint32_t depth4_of(const std::vector<HostBVH4Node>& n, int32_t i) {
  if (i < 0) return 0;
  uint32_t d = 0;
  for (uint32_t k = 0; k < n[i].num_children; k++) d = std::max(d, depth4_of(n, n[i].children[k]));
  return d + 1;
}
}  // namespace

void build_scene(HostScene& sc, std::vector<HostShape> shapes, std::vector<HostMaterial> mats, uint32_t bvh_kind, uint32_t num_bins) {
  if (bvh_kind != 2 && bvh_kind != 4) throw std::runtime_error("bvh_kind must be 2 or 4");
  for (size_t i = 0; i < shapes.size(); i++) shapes[i].source = (int32_t)i;
  // bvh.rs:376-394: infinite shapes are swapped to the front in encounter order
  std::vector<HostShape> finite;
  std::vector<Box> boxes;
  std::vector<V3> cents;
  uint32_t num_inf = 0;
  for (size_t i = 0; i < shapes.size(); i++) {
    Box b;
    if (shape_box(shapes[i], &b)) { finite.push_back(shapes[i]); boxes.push_back(b); cents.push_back(shape_centroid(shapes[i], b)); }
    else { std::swap(shapes[num_inf], shapes[i]); num_inf++; }
  }
  sc.bvh2.clear();
  sc.bvh2.push_back(HostBVH2Node{});
  sc.bvh2.push_back(HostBVH2Node{});   // pad: sibling pairs stay 64 B aligned (bvh.rs:108-109)
  if (!finite.empty()) {
    Builder bld(boxes, cents, sc.bvh2, num_bins);
    Box all = bld.range_box(0, finite.size());
    HostBVH2Node root = bld.subdivide(0, finite.size(), all);
    sc.bvh2[0] = root;
    for (size_t i = 0; i < finite.size(); i++) shapes[num_inf + i] = finite[bld.order[i]];
  }
  sc.shapes = std::move(shapes);
  sc.mats = std::move(mats);
  sc.num_inf = num_inf;
  sc.bvh_kind = bvh_kind;
  sc.depth2 = finite.empty() ? 0 : depth2_of(sc.bvh2, 0);
  sc.bvh4.clear();
  sc.depth4 = 0;
  if (bvh_kind == 4) {
    if (finite.empty()) throw std::runtime_error("BVH4: empty scene");
    Collapser col(sc.bvh2, sc.bvh4);
    col.run();
    sc.depth4 = depth4_of(sc.bvh4, 0);
  }
  sc.lights.clear();
  for (size_t i = 0; i < sc.shapes.size(); i++) if (sc.mats[sc.shapes[i].mat].emissive) sc.lights.push_back((uint32_t)i);
}

// ------------------------------------------------------------------ OBJ (obj_parser.ts:3-51)
static double js_parse_float(const char* b, const char* e) {
  std::string tmp(b, e);
  char* end = nullptr;
  double v = std::strtod(tmp.c_str(), &end);
  return end == tmp.c_str() ? std::numeric_limits<double>::quiet_NaN() : v;
}
std::vector<float> parse_obj_text(const char* text, size_t len, bool client_scale) {
  std::vector<double> verts;
  std::vector<long long> faces;
  const char* p = text;
  const char* end = text + len;
  while (p <= end) {
    const char* le = (const char*)memchr(p, '\n', (size_t)(end - p));
    if (!le) le = end;
    // fields are separated by SINGLE spaces (String.split(' '))
    const char* f[5]; const char* fe[5]; int nf = 0; int total = 0;
    const char* s = p;
    for (;;) {
      const char* sp = (const char*)memchr(s, ' ', (size_t)(le - s));
      const char* e = sp ? sp : le;
      if (nf < 5) { f[nf] = s; fe[nf] = e; nf++; }
      total++;
      if (!sp) break;
      s = sp + 1;
    }
    size_t l0 = (size_t)(fe[0] - f[0]);
    if (l0 == 1 && f[0][0] == 'v') {
      for (int k = 1; k <= 3; k++) verts.push_back(k < nf ? js_parse_float(f[k], fe[k]) : std::numeric_limits<double>::quiet_NaN());
    } else if (l0 == 1 && f[0][0] == 'f') {
      if (total != 4) throw std::runtime_error("Non-triangular face in OBJ file");
      for (int k = 1; k <= 3; k++) {
        std::string tmp(f[k], fe[k]);   // "a/b/c" -> parseInt(a)
        char* ep = nullptr;
        long long v = std::strtoll(tmp.c_str(), &ep, 10);
        faces.push_back(ep == tmp.c_str() ? -1 : v - 1);
      }
    }
    p = le + 1;
  }
  std::vector<float> out(faces.size() * 3);
  for (size_t i = 0; i < faces.size(); i++)
    for (int c = 0; c < 3; c++) {
      long long vi = faces[i] * 3 + c;
      out[i * 3 + c] = (faces[i] >= 0 && (size_t)vi < verts.size()) ? (float)verts[(size_t)vi] : std::numeric_limits<float>::quiet_NaN();
    }
  if (client_scale)   // index.ts:216-220
    for (size_t i = 0; i < out.size() / 3; i++) { out[i * 3] *= 8.0f; out[i * 3 + 1] *= 8.0f; out[i * 3 + 2] *= -8.0f; }
  return out;
}

// ------------------------------------------------------------------ flatten
static inline float bits_f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

void flatten_scene(const HostScene& sc, std::vector<DNode2>& n2, std::vector<DNode4>& n4, std::vector<DShape>& shp,
                   std::vector<DMaterial>& mats, std::vector<DLight>& lights) {
  n2.resize(sc.bvh2.size());
  for (size_t i = 0; i < sc.bvh2.size(); i++) {
    const HostBVH2Node& h = sc.bvh2[i];
    n2[i].a = make_float4(h.box.lo[0], h.box.lo[1], h.box.lo[2], h.box.hi[0]);
    n2[i].b = make_float4(h.box.hi[1], h.box.hi[2], bits_f(h.left_first), bits_f(h.count));
  }
  n4.resize(sc.bvh4.size());
  for (size_t i = 0; i < sc.bvh4.size(); i++) {
    const HostBVH4Node& h = sc.bvh4[i];
    DNode4 d{};
    d.x_min = make_float4(h.child[0].lo[0], h.child[1].lo[0], h.child[2].lo[0], h.child[3].lo[0]);
    d.y_min = make_float4(h.child[0].lo[1], h.child[1].lo[1], h.child[2].lo[1], h.child[3].lo[1]);
    d.z_min = make_float4(h.child[0].lo[2], h.child[1].lo[2], h.child[2].lo[2], h.child[3].lo[2]);
    d.x_max = make_float4(h.child[0].hi[0], h.child[1].hi[0], h.child[2].hi[0], h.child[3].hi[0]);
    d.y_max = make_float4(h.child[0].hi[1], h.child[1].hi[1], h.child[2].hi[1], h.child[3].hi[1]);
    d.z_max = make_float4(h.child[0].hi[2], h.child[1].hi[2], h.child[2].hi[2], h.child[3].hi[2]);
    d.children = make_int4(h.children[0], h.children[1], h.children[2], h.children[3]);
    d.num_children = h.num_children;
    n4[i] = d;
  }
  shp.resize(sc.shapes.size());
  for (size_t i = 0; i < sc.shapes.size(); i++) {
    const HostShape& s = sc.shapes[i];
    float meta = bits_f((uint32_t)s.type | (s.mat << 8));
    DShape d{};
    switch (s.type) {
      case SH_TRIANGLE: {
        V3 v0{s.p[0], s.p[1], s.p[2]}, v1{s.p[3], s.p[4], s.p[5]}, v2{s.p[6], s.p[7], s.p[8]};
        V3 n = crossv(sub(v1, v0), sub(v2, v0));   // triangle.rs:164
        V3 nn = normv(n);                          // triangle.rs:182
        d.q0 = make_float4(v0.x, v0.y, v0.z, meta);
        d.q1 = make_float4(v1.x, v1.y, v1.z, n.x);
        d.q2 = make_float4(v2.x, v2.y, v2.z, n.y);
        d.q3 = make_float4(n.z, nn.x, nn.y, nn.z);
        break;
      }
      case SH_PLANE: {
        V3 loc{s.p[0], s.p[1], s.p[2]}, nr{s.p[3], s.p[4], s.p[5]};
        d.q0 = make_float4(loc.x, loc.y, loc.z, meta);
        d.q1 = make_float4(nr.x, nr.y, nr.z, dotv(nr, loc));   // plane.rs:89
        break;
      }
      case SH_TORUS:
        d.q0 = make_float4(s.p[0], s.p[1], s.p[2], meta);
        d.q1 = make_float4(s.p[3], s.p[4], 0.0f, 0.0f);
        break;
      case SH_AARECT:
        d.q0 = make_float4(s.p[0], s.p[1], s.p[2], meta);
        d.q1 = make_float4(s.p[3], s.p[4], s.p[5], 0.0f);
        break;
      case SH_SPHERE: case SH_SQUARE:
        d.q0 = make_float4(s.p[0], s.p[1], s.p[2], meta);
        d.q1 = make_float4(s.p[3], 0.0f, 0.0f, 0.0f);
        break;
    }
    shp[i] = d;
  }
  mats.resize(sc.mats.size());
  for (size_t i = 0; i < sc.mats.size(); i++) {
    mats[i].c = make_float4(sc.mats[i].r, sc.mats[i].g, sc.mats[i].b, (float)sc.mats[i].kind);
    mats[i].p = make_float4(sc.mats[i].param, bits_f(sc.mats[i].tex), 0.0f, 0.0f);
  }
  lights.resize(sc.lights.size());
  for (size_t i = 0; i < sc.lights.size(); i++) {
    const HostShape& s = sc.shapes[sc.lights[i]];
    if (s.type != SH_TRIANGLE) throw std::runtime_error("only triangles can be area lights (ray.rs:96-105)");
    V3 v0{s.p[0], s.p[1], s.p[2]}, v1{s.p[3], s.p[4], s.p[5]}, v2{s.p[6], s.p[7], s.p[8]};
    float a = lenv(sub(v0, v1)), b = lenv(sub(v1, v2)), c = lenv(sub(v2, v0));   // triangle.rs:70-78
    float hs = (a + b + c) * 0.5f;
    float area = std::sqrt(hs * (hs - a) * (hs - b) * (hs - c));
    V3 nn = normv(crossv(sub(v1, v0), sub(v2, v0)));
    const HostMaterial& m = sc.mats[s.mat];
    lights[i].n_area = make_float4(nn.x, nn.y, nn.z, area);
    lights[i].intensity = make_float4(m.r, m.g, m.b, bits_f(sc.lights[i]));
  }
}

}  // namespace wpt
