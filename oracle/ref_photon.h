// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED. See ref_core.h.
// Restates src/math/empirical_pdf.rs and src/data/photon_tree.rs.
#pragma once
#include "ref_core.h"
#include <memory>

namespace ref {

// Mode-B bin contract (finding F9): photon weights are accumulated in
// unsigned fixed point with 40 fractional bits so that the per-node sums do not depend
// on insertion order (needed for a parallel / multi-GPU build); the f32 bin is then
// 1.0f + float(sum * 2^-40). Mode A keeps the reference's sequential f32 `+=`.
static const double PHOTON_FX_SCALE = 1099511627776.0;   // 2^40
static inline uint64_t photon_weight_fx(float w) {
  double s = (double)w * PHOTON_FX_SCALE;
  if (!(s > 0.0)) return 0;
  return (uint64_t)std::llrint(s);
}

struct EmpiricalPDF {   // empirical_pdf.rs:9-35
  std::vector<float> bins, cum_bins;
  std::vector<uint64_t> fx;
  bool fixed_point;
  bool has_updated_bins = true;
  EmpiricalPDF(size_t n, bool fixed) : bins(n, 1.0f), cum_bins(n, 0.0f), fx(fixed ? n : 0, 0), fixed_point(fixed) {}
  void set(size_t i, float v) { bins[i] = v; has_updated_bins = true; }
  void add(size_t i, float v) {   // empirical_pdf.rs:37-40
    if (fixed_point) fx[i] += photon_weight_fx(v);
    else bins[i] += v;
    has_updated_bins = true;
  }
  void recheck_cdf() {            // empirical_pdf.rs:79-93
    if (!has_updated_bins) return;
    if (fixed_point) for (size_t i = 0; i < bins.size(); i++) bins[i] = 1.0f + (float)((double)fx[i] * (1.0 / PHOTON_FX_SCALE));
    float bin_sum = 0.0f;
    for (float p : bins) bin_sum += p;
    cum_bins[0] = 0.0f;
    for (size_t i = 1; i < bins.size(); i++) cum_bins[i] = cum_bins[i - 1] + bins[i - 1] / bin_sum;
    has_updated_bins = false;
  }
  size_t sample(Rng& rng) {       // empirical_pdf.rs:43-61
    recheck_cdf();
    float r = rng.next();
    size_t low = 0, high = bins.size();
    while (low + 1 < high) {
      size_t mid = (low + high) / 2;
      if (cum_bins[mid] <= r) low = mid; else high = mid;
    }
    return low;
  }
  float bin_prob(size_t i) {      // empirical_pdf.rs:64-75
    recheck_cdf();
    if (i + 1 == cum_bins.size()) return 1.0f - cum_bins[i];
    return cum_bins[i + 1] - cum_bins[i];
  }
};

static const size_t MAX_PHOTONS_IN_CELL = 1024;   // photon_tree.rs:29

struct PhotonRec { size_t light; Vec3 loc; float w; };

// photon_tree.rs:235-251
static inline size_t octree_child(const AABB& b, Vec3 v, AABB* cb) {
  Vec3 c = b.center();
  size_t i = (v.x < c.x ? 0 : 4) + (v.y < c.y ? 0 : 2) + (v.z < c.z ? 0 : 1);
  *cb = AABB(v.x < c.x ? b.x_min : c.x, v.y < c.y ? b.y_min : c.y, v.z < c.z ? b.z_min : c.z,
             v.x < c.x ? c.x : b.x_max, v.y < c.y ? c.y : b.y_max, v.z < c.z ? c.z : b.z_max);
  return i;
}

struct Octree {   // photon_tree.rs:33-42
  bool is_node = false;
  EmpiricalPDF cdf;
  std::vector<Octree> children;
  std::vector<PhotonRec> values;
  Octree(size_t nl, bool fixed) : cdf(nl, fixed) {}

  // photon_tree.rs:165-196
  void insert(size_t nl, const AABB& self_bounds, size_t light, Vec3 loc, float w) {
    cdf.add(light, w);
    if (is_node) {
      AABB cb; size_t ci = octree_child(self_bounds, loc, &cb);
      children[ci].insert(nl, cb, light, loc, w);
      return;
    }
    values.push_back(PhotonRec{light, loc, w});
    if (values.size() > MAX_PHOTONS_IN_CELL) {
      bool fixed = cdf.fixed_point;
      std::vector<PhotonRec> old;
      old.swap(values);
      is_node = true;
      cdf = EmpiricalPDF(nl, fixed);
      children.assign(8, Octree(nl, fixed));
      for (auto& p : old) insert(nl, self_bounds, p.light, p.loc, p.w);
    }
  }
  // photon_tree.rs:201-211
  Octree* find_leaf(const AABB& b, size_t depth, Vec3 loc, AABB* out_b, size_t* out_depth) {
    if (is_node) { AABB cb; size_t ci = octree_child(b, loc, &cb); return children[ci].find_leaf(cb, depth + 1, loc, out_b, out_depth); }
    *out_b = b; *out_depth = depth;
    return this;
  }
  // photon_tree.rs:216-231
  EmpiricalPDF* find_node_cdf(const AABB& b, size_t depth, Vec3 loc) {
    if (is_node) {
      if (depth == 0) return &cdf;
      AABB cb; size_t ci = octree_child(b, loc, &cb);
      return children[ci].find_node_cdf(cb, depth - 1, loc);
    }
    return &cdf;
  }
  size_t count_nodes() const { size_t c = 1; for (auto& ch : children) c += ch.count_nodes(); return c; }
};

struct PhotonTree {   // photon_tree.rs:19-23, :48-56
  size_t num_lights;
  Octree root;
  float size = 1024.0f;
  PhotonTree(size_t nl, bool fixed) : num_lights(nl), root(nl, fixed) {}

  // photon_tree.rs:61-76 — quirk q4: the bounds check can never reject
  bool insert(size_t light, Vec3 loc, float w) {
    if (loc.x < -size && loc.x > size && loc.y < -size && loc.y > size && loc.z < -size && loc.z > size) return false;
    root.insert(num_lights, AABB(-size, -size, -size, size, size, size), light, loc, w);
    return true;
  }

  struct AxisW { float w, w_adj, off; };
  static AxisW axis_weight(float v, float c, float mn, float mx, float sz) {   // photon_tree.rs:90-124
    AxisW r;
    if (v > c) {
      float left_weight = (mx - (v - sz * 0.5f)) / sz;
      r.w = left_weight; r.w_adj = 1.0f - left_weight; r.off = 1.0f;
    } else {
      float right_weight = ((v + sz * 0.5f) - mn) / sz;
      r.w = right_weight; r.w_adj = 1.0f - right_weight; r.off = -1.0f;
    }
    return r;
  }

  // photon_tree.rs:80-159
  void sample(Rng& rng, Vec3 v, size_t* light, float* pdf_out) {
    if (v.x < -size || v.y < -size || v.z < -size || v.x > size || v.y > size || v.z > size) {
      *light = rng.next_in_range(0, num_lights);
      *pdf_out = 1.0f / (float)num_lights;
      return;
    }
    AABB self_bounds(-size, -size, -size, size, size, size);
    AABB b; size_t depth;
    root.find_leaf(self_bounds, 0, v, &b, &depth);
    Vec3 c = b.center();
    AxisW wx = axis_weight(v.x, c.x, b.x_min, b.x_max, b.x_size());
    AxisW wy = axis_weight(v.y, c.y, b.y_min, b.y_max, b.y_size());
    AxisW wz = axis_weight(v.z, c.z, b.z_min, b.z_max, b.z_size());
    // The reference asserts the weights are in [0,1] (photon_tree.rs:100,112,124).
    bool self_x = rng.next() <= wx.w;
    bool self_y = rng.next() <= wy.w;
    bool self_z = rng.next() <= wz.w;
    // v + a + b + c parses as ((v + a) + b) + c
    Vec3 sampled_v = v + (self_x ? Vec3() : wx.off * Vec3(b.x_size(), 0.0f, 0.0f));
    sampled_v = sampled_v + (self_y ? Vec3() : wy.off * Vec3(0.0f, b.y_size(), 0.0f));
    sampled_v = sampled_v + (self_z ? Vec3() : wz.off * Vec3(0.0f, 0.0f, b.z_size()));
    size_t res = root.find_node_cdf(self_bounds, depth, sampled_v)->sample(rng);
    float pdf = 0.0f;
    float ajx = b.x_size() * wx.off, ajy = b.y_size() * wy.off, ajz = b.z_size() * wz.off;
    auto P = [&](Vec3 q) { return root.find_node_cdf(self_bounds, depth, q)->bin_prob(res); };
    pdf += P(v) * wx.w * wy.w * wz.w;
    pdf += P(v + Vec3(ajx, 0.0f, 0.0f)) * wx.w_adj * wy.w * wz.w;
    pdf += P(v + Vec3(0.0f, ajy, 0.0f)) * wx.w * wy.w_adj * wz.w;
    pdf += P(v + Vec3(0.0f, 0.0f, ajz)) * wx.w * wy.w * wz.w_adj;
    pdf += P(v + Vec3(ajx, ajy, 0.0f)) * wx.w_adj * wy.w_adj * wz.w;
    pdf += P(v + Vec3(0.0f, ajy, ajz)) * wx.w * wy.w_adj * wz.w_adj;
    pdf += P(v + Vec3(ajx, 0.0f, ajz)) * wx.w_adj * wy.w * wz.w_adj;
    pdf += P(v + Vec3(ajx, ajy, ajz)) * wx.w_adj * wy.w_adj * wz.w_adj;
    *light = res;
    *pdf_out = pdf;
  }
};

}  // namespace ref
