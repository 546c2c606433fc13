// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED. See ref_core.h.
// Restates src/graphics/bvh.rs (binned BVH2 build) and src/graphics/bvh4.rs
// (tree-cut collapse to a 4-wide BVH).
#pragma once
#include "ref_shapes.h"

namespace ref {

// bvh.rs:14-20
struct BVHNode {
  AABB bounds;
  uint32_t left_first = 0;
  uint32_t count = 0;
  bool is_leaf() const { return count > 0; }
};

struct ShapeRep {   // bvh.rs:86-90
  Shape shape;
  Vec3 location;
  AABB bounds;
};

struct BinResult {  // bvh.rs:440-476
  std::vector<std::vector<ShapeRep>> bins;
  explicit BinResult(size_t nb) : bins(nb) {}
  void clear() { for (auto& b : bins) b.clear(); }
  size_t num_bins() const { return bins.size(); }
  void write_to(ShapeRep* dst) const {
    size_t i = 0;
    for (auto& b : bins) for (auto& v : b) dst[i++] = v;
  }
};

static inline bool reps_aabb(const ShapeRep* s, size_t n, AABB* out) {   // bvh.rs:397-407
  if (n == 0) return false;
  AABB r = s[0].bounds;
  for (size_t i = 1; i < n; i++) r = r.join(s[i].bounds);
  *out = r;
  return true;
}
static inline bool bin_aabb(const std::vector<ShapeRep>& b, AABB* out) { return reps_aabb(b.data(), b.size(), out); }
static inline AABB join_maybe(const AABB& a, const std::vector<ShapeRep>& b) {   // aabb.rs:103-109
  AABB o;
  if (bin_aabb(b, &o)) return a.join(o);
  return a;
}

// bvh.rs:412-437
static inline bool bin_shapes(const ShapeRep* xs, size_t n, int axis, BinResult& dst) {
  auto f = [axis](const ShapeRep& s) { return axis == 0 ? s.location.x : (axis == 1 ? s.location.y : s.location.z); };
  float min_v = f(xs[0]), max_v = f(xs[0]);
  for (size_t i = 1; i < n; i++) { float v = f(xs[i]); min_v = fmin_(min_v, v); max_v = fmax_(max_v, v); }
  if (min_v == max_v) return false;
  size_t nb = dst.num_bins();
  dst.clear();
  float segment_width = (max_v - min_v) / (float)nb;
  for (size_t i = 0; i < n; i++) {
    float v = f(xs[i]);
    float q = std::floor((v - min_v) / segment_width);
    size_t id = q >= 0.0f ? (size_t)q : 0;     // Rust `as usize` saturates
    id = std::min(id, nb - 1);
    dst.bins[id].push_back(xs[i]);
  }
  return true;
}

// bvh.rs:309-370
static inline bool split_axis(const ShapeRep* shapes, size_t n, int axis, BinResult& bins, AABB* l_out, AABB* r_out, size_t* idx) {
  size_t num_bins = bins.num_bins();
  if (n <= 1) return false;
  if (!bin_shapes(shapes, n, axis, bins)) return false;
  size_t l = 0, r = num_bins - 1;
  AABB l_aabb, r_aabb;
  bin_aabb(bins.bins[l], &l_aabb);
  bin_aabb(bins.bins[r], &r_aabb);
  size_t l_cnt = bins.bins[l].size(), r_cnt = bins.bins[r].size();
  AABB ln_aabb = join_maybe(l_aabb, bins.bins[l + 1]);
  AABB rn_aabb = join_maybe(r_aabb, bins.bins[r - 1]);
  size_t ln_cnt = l_cnt + bins.bins[l + 1].size();
  size_t rn_cnt = r_cnt + bins.bins[r - 1].size();
  while (l + 1 < r) {
    if ((ln_aabb.surface() * (float)ln_cnt + r_aabb.surface() * (float)r_cnt) <
        (l_aabb.surface() * (float)l_cnt + rn_aabb.surface() * (float)rn_cnt)) {
      l += 1; l_aabb = ln_aabb; l_cnt = ln_cnt;
      if (l + 1 < r) { ln_aabb = join_maybe(l_aabb, bins.bins[l + 1]); ln_cnt = l_cnt + bins.bins[l + 1].size(); }
    } else {
      r -= 1; r_aabb = rn_aabb; r_cnt = rn_cnt;
      if (l + 1 < r) { rn_aabb = join_maybe(r_aabb, bins.bins[r - 1]); rn_cnt = r_cnt + bins.bins[r - 1].size(); }
    }
  }
  *l_out = l_aabb; *r_out = r_aabb; *idx = l_cnt;
  return true;
}

// bvh.rs:282-303 — longest axis of the box handed down; ties prefer z, then y
static inline bool split_longest_axis(const ShapeRep* shapes, size_t n, const AABB& parent, BinResult& bins, AABB* l, AABB* r, size_t* idx) {
  float xs = parent.x_max - parent.x_min, ys = parent.y_max - parent.y_min, zs = parent.z_max - parent.z_min;
  int axis;
  if (xs > ys) axis = (xs > zs) ? 0 : 2;
  else if (ys > zs) axis = 1;
  else axis = 2;
  return split_axis(shapes, n, axis, bins, l, r, idx);
}

// bvh.rs:215-277 (subdivide + split)
static inline BVHNode subdivide(std::vector<BVHNode>& dst, std::vector<ShapeRep>& shapes, size_t offset, size_t length, const AABB& parent_aabb, BinResult& bins) {
  ShapeRep* sl = shapes.data() + offset;
  bool do_split = false;
  AABB l_aabb, r_aabb, leaf_aabb;
  size_t split_index = 0;
  if (length <= 1) {
    reps_aabb(sl, length, &leaf_aabb);
  } else if (split_longest_axis(sl, length, parent_aabb, bins, &l_aabb, &r_aabb, &split_index)) {
    float utility = l_aabb.surface() * (float)split_index + r_aabb.surface() * (float)(length - split_index);
    AABB joined = l_aabb.join(r_aabb);
    float parent_utility = joined.surface() * (float)length;
    if (utility < parent_utility) { bins.write_to(sl); do_split = true; }
    else leaf_aabb = joined;
  } else {
    reps_aabb(sl, length, &leaf_aabb);
  }
  BVHNode node;
  if (do_split) {
    size_t left_id = dst.size();
    dst.push_back(BVHNode());
    dst.push_back(BVHNode());
    BVHNode ln = subdivide(dst, shapes, offset, split_index, l_aabb, bins);
    dst[left_id] = ln;
    BVHNode rn = subdivide(dst, shapes, offset + split_index, length - split_index, r_aabb, bins);
    dst[left_id + 1] = rn;
    node.bounds = l_aabb.join(r_aabb);
    node.left_first = (uint32_t)left_id;
    node.count = 0;
  } else {
    node.bounds = leaf_aabb;
    node.left_first = (uint32_t)offset;
    node.count = (uint32_t)length;
  }
  return node;
}

// bvh.rs:103-125 + shape_reps :376-394. Reorders `shapes`: infinite first, then BVH order.
static inline size_t build_bvh(std::vector<Shape>& shapes, size_t num_bins, std::vector<BVHNode>& dst) {
  size_t num_infinite = 0;
  std::vector<ShapeRep> reps;
  for (size_t i = 0; i < shapes.size(); i++) {
    AABB b; Vec3 loc;
    if (shapes[i].aabb(&b) && shapes[i].centroid(&loc)) {
      reps.push_back(ShapeRep{shapes[i], loc, b});
    } else {
      std::swap(shapes[num_infinite], shapes[i]);
      num_infinite++;
    }
  }
  dst.clear();
  dst.push_back(BVHNode());
  dst.push_back(BVHNode());   // pad so that sibling pairs share a 64 B line (bvh.rs:108-109)
  if (reps.empty()) return num_infinite;
  BinResult bins(num_bins);
  AABB all;
  reps_aabb(reps.data(), reps.size(), &all);
  BVHNode root = subdivide(dst, reps, 0, reps.size(), all, bins);
  dst[0] = root;
  for (size_t i = 0; i < reps.size(); i++) shapes[i + num_infinite] = reps[i].shape;
  return num_infinite;
}

static inline uint32_t bvh_depth(const std::vector<BVHNode>& nodes, size_t i = 0) {   // bvh.rs:202-210
  if (nodes[i].count != 0) return 0;
  return 1 + std::max(bvh_depth(nodes, nodes[i].left_first), bvh_depth(nodes, nodes[i].left_first + 1));
}
static inline size_t bvh_node_count(const std::vector<BVHNode>& nodes, size_t i = 0) {   // bvh.rs:67-79
  if (nodes[i].is_leaf()) return 1;
  return 1 + bvh_node_count(nodes, nodes[i].left_first) + bvh_node_count(nodes, nodes[i].left_first + 1);
}
// bvh.rs:128-194 (never called by the reference; used by the tests)
static inline bool verify_bvh_bounds(const std::vector<Shape>& shapes, size_t num_inf, const std::vector<BVHNode>& bvh, size_t i) {
  const BVHNode& n = bvh[i];
  if (n.count == 0) {
    if (!verify_bvh_bounds(shapes, num_inf, bvh, n.left_first)) return false;
    if (!verify_bvh_bounds(shapes, num_inf, bvh, n.left_first + 1)) return false;
    return n.bounds.contains(bvh[n.left_first].bounds.join(bvh[n.left_first + 1].bounds));
  }
  for (size_t k = num_inf + n.left_first; k < num_inf + n.left_first + n.count; k++) {
    AABB b;
    if (!shapes[k].aabb(&b) || !n.bounds.contains(b)) return false;
  }
  return true;
}
static inline void verify_bvh_contains(std::vector<char>& seen, const std::vector<BVHNode>& bvh, size_t i) {
  if (bvh[i].count == 0) { verify_bvh_contains(seen, bvh, bvh[i].left_first); verify_bvh_contains(seen, bvh, bvh[i].left_first + 1); }
  else for (uint32_t k = bvh[i].left_first; k < bvh[i].left_first + bvh[i].count; k++) seen[k] = 1;
}
static inline bool verify_bvh(const std::vector<Shape>& shapes, size_t num_inf, const std::vector<BVHNode>& bvh) {
  if (shapes.size() == num_inf) return true;
  bool a = verify_bvh_bounds(shapes, num_inf, bvh, 0);
  std::vector<char> seen(shapes.size() - num_inf, 0);
  verify_bvh_contains(seen, bvh, 0);
  for (char c : seen) a = a && c;
  return a;
}

// ================================================================ BVH4 (bvh4.rs)
// bvh4.rs:17-26. child_bounds is kept as 4 AABBs (the SoA f32x4 lanes of AABBx4).
struct BVHNode4 {
  AABB child_bounds[4];
  int32_t children[4] = {INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN};
  uint32_t num_children = 0;
  AABB extract_hull(size_t n) const {   // aabb.rs:226-233
    AABB h = child_bounds[0];
    for (size_t i = 1; i < n; i++) h = h.join(child_bounds[i]);
    return h;
  }
};

// DEVIATION (finding F7): the reference encodes `count << 27` but decodes
// `& 0x3`, silently dropping shapes of leaves with more than 3 primitives. Both this
// oracle and the product decode the documented 4 bits (`& 0xF`); leaves with more than
// 15 shapes cannot be encoded and are reported as an error.
static const uint32_t BVH4_COUNT_MASK = 0xF;
static inline uint32_t bvh4_leaf_count(int32_t code) { return ((uint32_t)code >> 27) & BVH4_COUNT_MASK; }
static inline uint32_t bvh4_leaf_first(int32_t code) { return (uint32_t)code & 0x7FFFFFFu; }

typedef std::vector<std::vector<float>> Bvh4Memo;   // empty vector == None

// bvh4.rs:244-281
static inline float r_cost(Bvh4Memo& memo, const std::vector<BVHNode>& bvh, size_t node_i, size_t cutsize) {
  const float t_cost = 1.0f;
  const size_t max_childs = 4;
  if (bvh[node_i].is_leaf()) return t_cost;
  size_t li = bvh[node_i].left_first, ri = li + 1;
  if (memo[node_i].empty()) {
    std::vector<float> cost(max_childs, INF_F);
    for (size_t t = 2; t <= max_childs; t++) {
      for (size_t i = 1; i < t; i++) {
        float r = r_cost(memo, bvh, li, i) + r_cost(memo, bvh, ri, t - i);
        cost[t - 1] = fmin_(cost[t - 1], r);
      }
      cost[0] = fmin_(cost[0], t_cost + cost[t - 1]);
    }
    memo[node_i] = cost;
  }
  const std::vector<float>& m = memo[node_i];
  if (cutsize == 0) return 0.0f;
  float cut_min = m[0];
  for (size_t i = 1; i < cutsize; i++) cut_min = fmin_(cut_min, m[i]);
  return cut_min;
}
// bvh4.rs:228-240
static inline float node_flat_cost(const Bvh4Memo& memo, const std::vector<BVHNode>& bvh, size_t node_i, size_t cutsize) {
  if (bvh[node_i].is_leaf()) return 1.0f;
  if (!memo[node_i].empty()) {
    float cut_min = memo[node_i][0];
    for (size_t i = 1; i < cutsize; i++) cut_min = fmin_(cut_min, memo[node_i][i]);
    return cut_min;
  }
  return INF_F;
}
// bvh4.rs:189-205
static inline size_t find_t(const std::vector<BVHNode>& bvh, const Bvh4Memo& memo, size_t node_i, size_t cutsize) {
  if (bvh[node_i].is_leaf()) return 1;
  if (memo[node_i].empty()) throw std::runtime_error("INVALID T");
  const std::vector<float>& m = memo[node_i];
  size_t t_min = 1; float t_min_val = m[0];
  for (size_t t = 2; t <= cutsize; t++) if (m[t - 1] < t_min_val) { t_min = t; t_min_val = m[t - 1]; }
  return t_min;
}
// bvh4.rs:210-224
static inline size_t find_i(const std::vector<BVHNode>& bvh, const Bvh4Memo& memo, size_t li, size_t ri, size_t t) {
  size_t i_min = 1;
  float i_min_val = node_flat_cost(memo, bvh, li, 1) + node_flat_cost(memo, bvh, ri, t - 1);
  for (size_t i = 2; i < t; i++) {
    float v = node_flat_cost(memo, bvh, li, i) + node_flat_cost(memo, bvh, ri, t - i);
    if (v < i_min_val) { i_min = i; i_min_val = v; }
  }
  return i_min;
}
typedef std::vector<std::pair<AABB, int32_t>> ChildList;
// bvh4.rs:127-185
static inline ChildList collapse_with(std::vector<BVHNode4>& dst, const std::vector<BVHNode>& bvh, const Bvh4Memo& memo, size_t node_i, size_t cutsize) {
  if (bvh[node_i].is_leaf()) {
    if (bvh[node_i].count > BVH4_COUNT_MASK) throw std::runtime_error("BVH4: leaf with more than 15 shapes cannot be encoded");
    uint32_t code = 0x80000000u | (bvh[node_i].count << 27) | bvh[node_i].left_first;
    return ChildList{{bvh[node_i].bounds, (int32_t)code}};
  }
  size_t li = bvh[node_i].left_first, ri = li + 1;
  size_t t = find_t(bvh, memo, node_i, cutsize);
  if (t == 1) {
    size_t index = dst.size();
    dst.push_back(BVHNode4());
    size_t i_min = find_i(bvh, memo, li, ri, 4);
    ChildList lcs = collapse_with(dst, bvh, memo, li, i_min);
    ChildList rcs = collapse_with(dst, bvh, memo, ri, 4 - i_min);
    BVHNode4 n;
    size_t j = 0;
    for (auto& e : lcs) { n.children[j] = e.second; n.child_bounds[j] = e.first; j++; }
    for (auto& e : rcs) { n.children[j] = e.second; n.child_bounds[j] = e.first; j++; }
    n.num_children = (uint32_t)(lcs.size() + rcs.size());
    dst[index] = n;
    return ChildList{{n.extract_hull(n.num_children), (int32_t)index}};
  }
  size_t i_min = find_i(bvh, memo, li, ri, t);
  ChildList c1 = collapse_with(dst, bvh, memo, li, i_min);
  ChildList c2 = collapse_with(dst, bvh, memo, ri, t - i_min);
  c1.insert(c1.end(), c2.begin(), c2.end());
  return c1;
}
// bvh4.rs:37-70
static inline std::vector<BVHNode4> collapse_bvh4(const std::vector<BVHNode>& bvh2) {
  Bvh4Memo memo(bvh2.size());
  r_cost(memo, bvh2, 0, 4);
  std::vector<BVHNode4> dst;
  ChildList res = collapse_with(dst, bvh2, memo, 0, 4);
  if (res.size() > 1) {
    dst.clear();
    dst.push_back(BVHNode4());
    ChildList res2 = collapse_with(dst, bvh2, memo, 0, 4);
    BVHNode4 n;
    n.children[0] = n.children[1] = n.children[2] = n.children[3] = 0;   // bvh4.rs:55
    for (size_t i = 0; i < res2.size(); i++) { n.child_bounds[i] = res2[i].first; n.children[i] = res2[i].second; }
    n.num_children = (uint32_t)res2.size();
    dst[0] = n;
  } else if (res[0].second != 0) {
    throw std::runtime_error("BVH4: root is a leaf (bvh4.rs:67 assert)");
  }
  return dst;
}
static inline size_t bvh4_node_count(const std::vector<BVHNode4>& b, int32_t i = 0) {   // bvh4.rs:79-89
  if (i < 0) return 1;
  size_t c = 1;
  for (uint32_t j = 0; j < b[i].num_children; j++) c += bvh4_node_count(b, b[i].children[j]);
  return c;
}
static inline size_t bvh4_depth(const std::vector<BVHNode4>& b, int32_t i = 0) {   // bvh4.rs:99-109
  if (i < 0) return 0;
  size_t d = 0;
  for (uint32_t j = 0; j < b[i].num_children; j++) d = std::max(d, bvh4_depth(b, b[i].children[j]));
  return d + 1;
}
// bvh4.rs:300-376
static inline bool verify_bvh4_bounds(const std::vector<Shape>& shapes, size_t num_inf, const std::vector<BVHNode4>& b, const AABB& bounds, int32_t i) {
  if (i >= 0) {
    const BVHNode4& n = b[i];
    if (n.num_children > 4) return false;
    for (uint32_t j = 0; j < n.num_children; j++)
      if (!verify_bvh4_bounds(shapes, num_inf, b, n.child_bounds[j], n.children[j])) return false;
    return true;
  }
  uint32_t cnt = bvh4_leaf_count(i), first = bvh4_leaf_first(i);
  for (size_t k = num_inf + first; k < num_inf + first + cnt; k++) {
    AABB sb;
    if (!shapes[k].aabb(&sb) || !bounds.contains(sb)) return false;
  }
  return true;
}
static inline void verify_bvh4_contains(std::vector<char>& seen, const std::vector<BVHNode4>& b, int32_t i) {
  if (i >= 0) { for (uint32_t j = 0; j < b[i].num_children; j++) verify_bvh4_contains(seen, b, b[i].children[j]); }
  else { uint32_t cnt = bvh4_leaf_count(i), first = bvh4_leaf_first(i); for (uint32_t k = 0; k < cnt; k++) seen[first + k] = 1; }
}
static inline bool verify_bvh4(const std::vector<Shape>& shapes, size_t num_inf, const std::vector<BVHNode4>& b) {
  AABB self = b[0].extract_hull(b[0].num_children);
  bool a = verify_bvh4_bounds(shapes, num_inf, b, self, 0);
  std::vector<char> seen(shapes.size() - num_inf, 0);
  verify_bvh4_contains(seen, b, 0);
  for (char c : seen) a = a && c;
  return a;
}

}  // namespace ref
