// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED. See ref_core.h.
// Restates src/graphics/scene.rs (trace_g, traverse_bvh*, shadow_ray), src/scenes.rs,
// src_ts/client/obj_parser.ts and the mesh transforms of src_ts/client/index.ts:216-220
// and src/wasm_interface.rs:297-313.
#pragma once
#include "ref_bvh.h"
#include <cstdio>
#include <cstdlib>

namespace ref {

struct TraceCounters {
  uint64_t rays = 0, node_visits = 0, prim_tests = 0;
};

struct GHit { bool some = false; float dis = 0; size_t shape = 0; };

// primitive-test counter (P in SURVEY 8d); thread-local so that the threaded CPU baseline
// does not race on it
inline uint64_t& tl_prim_tests() { static thread_local uint64_t c = 0; return c; }

struct Scene {
  Color3 background;
  std::vector<size_t> lights;     // LightEnum::Area(shape index) — scene.rs:62-66
  std::vector<Shape> shapes;
  int bvh_kind = 2;               // 2, 4 or 0 (none)
  size_t num_inf = 0;
  std::vector<BVHNode> bvh2;
  std::vector<BVHNode4> bvh4;

  // scene.rs:43-69. `use_bvh4` is the one knob the reference hard-wires to false (F6).
  void build(Color3 bg, std::vector<Shape> shp, bool use_bvh4, size_t num_bins = 16) {
    background = bg;
    shapes = std::move(shp);
    for (size_t i = 0; i < shapes.size(); i++) shapes[i].source_index = (int)i;
    num_inf = build_bvh(shapes, num_bins, bvh2);
    bvh_kind = 2;
    if (use_bvh4) {
      bvh4 = collapse_bvh4(bvh2);
      if (!verify_bvh4(shapes, num_inf, bvh4)) throw std::runtime_error("WHAT");   // scene.rs:84-87
      bvh_kind = 4;
    }
    lights.clear();
    for (size_t i = 0; i < shapes.size(); i++) if (shapes[i].is_emissive()) lights.push_back(i);
  }

  // scene.rs:426-445
  GHit trace_shapes(const Ray& ray, size_t first, size_t count) const {
    GHit best;
    for (size_t i = 0; i < count; i++) {
      float nd;
      tl_prim_tests()++;
      if (shapes[first + i].trace_simple(ray, &nd)) {
        if (best.some) { if (0.0f < nd && nd < best.dis) { best.dis = nd; best.shape = i; } }
        else { best.some = true; best.dis = nd; best.shape = i; }
      }
    }
    return best;
  }
  // scene.rs:450-472
  GHit trace_shapes_md(const Ray& ray, size_t first, size_t count, float max_dis) const {
    GHit best;
    for (size_t i = 0; i < count; i++) {
      float nd;
      tl_prim_tests()++;
      if (shapes[first + i].trace_simple(ray, &nd)) {
        if (nd <= max_dis) {
          if (best.some) { if (0.0f < nd && nd < best.dis) { best.dis = nd; best.shape = i; } }
          else { best.some = true; best.dis = nd; best.shape = i; }
        }
      }
    }
    return best;
  }
  // scene.rs:393-403
  static bool aabb_distance(const Ray& ray, const AABB& b, float max_dis, float* d) {
    float h;
    if (b.hit(ray, &h) && h < max_dis) { *d = h; return true; }
    return false;
  }
  // scene.rs:191-212
  size_t traverse_bvh_guarded(const Ray& ray, size_t node_i, float max_dis, GHit* res) const {
    float h;
    if (bvh2[node_i].bounds.hit(ray, &h) && h < max_dis) return traverse_bvh(ray, node_i, max_dis, res) + 1;
    res->some = false;
    return 1;
  }
  // scene.rs:218-288
  size_t traverse_bvh(const Ray& ray, size_t node_i, float max_dis, GHit* res) const {
    const BVHNode& node = bvh2[node_i];
    if (node.count != 0) {
      GHit h = trace_shapes_md(ray, num_inf + node.left_first, node.count, max_dis);
      if (h.some) { res->some = true; res->dis = h.dis; res->shape = num_inf + node.left_first + h.shape; }
      else res->some = false;
      return 1;
    }
    size_t li = node.left_first;
    float left_dis, right_dis;
    if (aabb_distance(ray, bvh2[li].bounds, max_dis, &left_dis)) {
      if (aabb_distance(ray, bvh2[li + 1].bounds, max_dis, &right_dis)) {
        if (left_dis < right_dis) {
          GHit tl; size_t ld = traverse_bvh(ray, li, max_dis, &tl);
          if (tl.some) {
            if (tl.dis < right_dis) { *res = tl; return 1 + ld; }
            GHit tr; size_t rd = traverse_bvh(ray, li + 1, tl.dis, &tr);
            *res = tr.some ? tr : tl;
            return 1 + ld + rd;
          }
          GHit tr; size_t rd = traverse_bvh(ray, li + 1, max_dis, &tr);
          *res = tr;
          return 1 + ld + rd;
        } else {
          GHit tr; size_t rd = traverse_bvh(ray, li + 1, max_dis, &tr);
          if (tr.some) {
            if (tr.dis < left_dis) { *res = tr; return 1 + rd; }
            GHit tl; size_t ld = traverse_bvh(ray, li, tr.dis, &tl);
            *res = tl.some ? tl : tr;
            return 1 + ld + rd;
          }
          GHit tl; size_t ld = traverse_bvh(ray, li, max_dis, &tl);
          *res = tl;
          return 1 + ld + rd;
        }
      }
      size_t ld = traverse_bvh(ray, li, max_dis, res);
      return ld + 1;
    }
    size_t rd = traverse_bvh_guarded(ray, li + 1, max_dis, res);
    return rd + 1;
  }

  // scene.rs:346-388 — the exact compare-and-swap sequence (not stable for n == 4)
  static void sort_small(std::pair<int32_t, float>* a, size_t n) {
    if (n == 2) {
      if (a[1].second < a[0].second) std::swap(a[0], a[1]);
    } else if (n == 3) {
      if (a[1].second < a[0].second) std::swap(a[0], a[1]);
      if (a[2].second < a[1].second) std::swap(a[1], a[2]);
      if (a[1].second < a[0].second) std::swap(a[0], a[1]);
    } else if (n == 4) {
      if (a[1].second < a[0].second) std::swap(a[0], a[1]);
      if (a[3].second < a[2].second) std::swap(a[2], a[3]);
      if (a[0].second < a[2].second) {
        if (a[2].second < a[1].second) {
          std::swap(a[1], a[2]);
          if (a[3].second < a[2].second) std::swap(a[2], a[3]);
        }
      } else {
        std::swap(a[0], a[2]);
        std::swap(a[1], a[2]);
        if (a[3].second < a[1].second) { std::swap(a[1], a[3]); std::swap(a[2], a[3]); }
        else if (a[3].second < a[2].second) std::swap(a[2], a[3]);
      }
    }
  }
  // scene.rs:292-342
  size_t traverse_bvh4(const Ray& ray, int32_t node_i, float max_dis, GHit* res) const {
    if (node_i < 0) {
      uint32_t cnt = bvh4_leaf_count(node_i), first = bvh4_leaf_first(node_i);
      GHit h = trace_shapes_md(ray, num_inf + first, cnt, max_dis);
      if (h.some) { res->some = true; res->dis = h.dis; res->shape = num_inf + first + h.shape; }
      else res->some = false;
      return 1;
    }
    const BVHNode4& node = bvh4[node_i];
    size_t nc = node.num_children;
    std::pair<int32_t, float> ch[4] = {{0, INF_F}, {0, INF_F}, {0, INF_F}, {0, INF_F}};
    for (size_t i = 0; i < nc; i++) ch[i] = {node.children[i], node.child_bounds[i].hit_x4_lane(ray)};
    sort_small(ch, nc);
    size_t traversed = 1;
    res->some = false;
    for (size_t i = 0; i < nc; i++) {
      if (ch[i].second > max_dis) return traversed;
      if (ch[i].second >= 0.0f) {
        GHit r2; size_t nt2 = traverse_bvh4(ray, ch[i].first, max_dis, &r2);
        if (r2.some) { max_dis = r2.dis; *res = r2; }
        traversed += nt2;
      }
    }
    return traversed;
  }

  // scene.rs:406-422 — tie goes to b
  static GHit closest(const GHit& a, const GHit& b) {
    if (a.some) { if (b.some) return a.dis < b.dis ? a : b; return a; }
    return b;
  }
  // scene.rs:162-184
  size_t trace_g(const Ray& ray, GHit* out) const {
    if (bvh_kind == 0) { *out = trace_shapes(ray, 0, shapes.size()); return 0; }
    GHit h1 = trace_shapes(ray, 0, num_inf);
    float md = h1.some ? h1.dis : INF_F;
    GHit h2; size_t d;
    if (bvh_kind == 2) d = traverse_bvh_guarded(ray, 0, md, &h2);
    else d = traverse_bvh4(ray, 0, md, &h2);
    *out = h1.some ? closest(h1, h2) : h2;
    return d;
  }
  // scene.rs:137-144 — the winner is intersected a second time for normal + material
  size_t trace(const Ray& ray, bool* some, Hit* hit, size_t* shape_id = nullptr) const {
    GHit g; size_t d = trace_g(ray, &g);
    if (g.some) { *some = shapes[g.shape].trace(ray, hit); if (shape_id) *shape_id = g.shape; }
    else *some = false;
    return d;
  }
  // scene.rs:104-133 — quirk q5: `dis` is measured from the offset origin
  size_t shadow_ray(Vec3 p, Vec3 point_on_shape, size_t light_shape, bool* occluded) const {
    Vec3 dir = point_on_shape - p;
    float dir_len = len(dir);
    dir = dir / dir_len;
    Ray ray(p + dir * EPSILON, dir);
    GHit g; size_t d = trace_g(ray, &g);
    if (g.some && g.dis < dir_len) *occluded = (g.shape != light_shape);
    else *occluded = false;
    return d;
  }
};

// ---------------------------------------------------------------- scenes.rs
static inline void museum_lights(std::vector<Shape>& dst, float x, float y, Vec3 color) {   // scenes.rs:54-68
  Vec3 lc1(x - 1.0f, 0.0f, y + 2.8f), lc2(x + 1.0f, 0.0f, y + 2.8f), lc3(x + 1.0f, 1.0f, y + 2.5f), lc4(x - 1.0f, 1.0f, y + 2.5f);
  dst.push_back(Shape::triangle(lc3, lc2, lc1, Material::emit(color)));
  dst.push_back(Shape::triangle(lc4, lc3, lc1, Material::emit(color)));
  lc1 = Vec3(x - 1.0f, 0.0f, y - 2.8f); lc2 = Vec3(x + 1.0f, 0.0f, y - 2.8f); lc3 = Vec3(x + 1.0f, 1.0f, y - 2.5f); lc4 = Vec3(x - 1.0f, 1.0f, y - 2.5f);
  dst.push_back(Shape::triangle(lc3, lc2, lc1, Material::emit(color)));
  dst.push_back(Shape::triangle(lc4, lc3, lc1, Material::emit(color)));
}
// scenes.rs:15-52. `color_order_out` (optional) receives, per row, the indices into the
// initial colour list — the KAT cross-checked against banner.png.
static inline std::vector<Shape> museum_shapes(std::vector<int>* color_order_out = nullptr, uint32_t* rng_state_out = nullptr) {
  std::vector<Shape> shapes;
  shapes.push_back(Shape::plane(Vec3(0.0f, -1.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), Material::diffuse(Color3(0.7f, 0.7f, 0.7f))));
  const float xs[9] = {-16.0f, -12.0f, -8.0f, -4.0f, 0.0f, 4.0f, 8.0f, 12.0f, 16.0f};
  std::vector<std::pair<Color3, int>> colors = {
      {Color3(1.0f, 0.3f, 0.3f), 0}, {Color3(0.0f, 1.0f, 1.0f), 1}, {Color3(0.3f, 0.3f, 1.0f), 2}, {Color3(1.0f, 0.0f, 0.0f), 3},
      {Color3(0.0f, 1.0f, 0.0f), 4}, {Color3(0.0f, 0.0f, 1.0f), 5}, {Color3(1.0f, 0.0f, 1.0f), 6}, {Color3(1.0f, 1.0f, 0.0f), 7},
      {Color3(0.3f, 1.0f, 0.3f), 8}};
  Rng rng;
  rng.next();
  rng.next();
  const float rows[3] = {-7.5f, 0.0f, 7.5f};
  for (float y : rows) {
    for (int i = 0; i < 9; i++) {
      shapes.push_back(Shape::torus(Vec3(xs[i], -0.5f, y), 1.3f, 0.3f, Material::diffuse(Color3(1.0f, 1.0f, 1.0f))));
      museum_lights(shapes, xs[i], y, colors[i].first.to_vec3() * 2.5f);
      if (color_order_out) color_order_out->push_back(colors[i].second);
    }
    rng.shuffle(colors);
  }
  if (rng_state_out) *rng_state_out = rng.state;
  const float ws[8] = {-14.0f, -10.0f, -6.0f, -2.0f, 2.0f, 6.0f, 10.0f, 14.0f};
  for (float x : ws) shapes.push_back(Shape::aarect(x - 0.1f, x + 0.1f, -1.0f, 2.0f, -20.0f, 20.0f, Material::diffuse(Color3(0.7f, 0.7f, 0.7f))));
  shapes.push_back(Shape::aarect(-20.0f, 20.0f, -1.0f, 2.0f, 3.75f - 0.1f, 3.75f + 0.1f, Material::diffuse(Color3(0.7f, 0.7f, 0.7f))));
  shapes.push_back(Shape::aarect(-20.0f, 20.0f, -1.0f, 2.0f, -3.75f - 0.1f, -3.75f + 0.1f, Material::diffuse(Color3(0.7f, 0.7f, 0.7f))));
  return shapes;
}
// scenes.rs:75-111; `mesh_tris` = the Mesh::Triangled list (may be null: mesh not loaded)
static inline std::vector<Shape> bunny_shapes(const std::vector<Shape>* mesh_tris) {
  std::vector<Shape> shapes;
  shapes.push_back(Shape::plane(Vec3(0.0f, -1.0f, 0.0f), Vec3(0.0f, 1.0f, 0.0f), Material::diffuse(Color3(1.0f, 1.0f, 1.0f))));
  shapes.push_back(Shape::plane(Vec3(0.0f, 0.0f, 13.0f), Vec3(0.0f, 0.0f, -1.0f), Material::diffuse(Color3(0.8f, 1.0f, 0.8f))));
  if (mesh_tris) for (auto& t : *mesh_tris) shapes.push_back(t);
  Vec3 lc1(-1.0f, 7.0f, 0.0f), lc2(1.0f, 7.0f, 0.0f), lc3(1.0f, 7.0f, 2.0f), lc4(-1.0f, 7.0f, 2.0f);
  shapes.push_back(Shape::triangle(lc3, lc2, lc1, Material::emit(Vec3(16.0f, 16.0f, 16.0f))));
  shapes.push_back(Shape::triangle(lc4, lc3, lc1, Material::emit(Vec3(16.0f, 16.0f, 16.0f))));
  return shapes;
}
// wasm_interface.rs:297-313 — Preload vertices (9 floats per triangle) -> Triangled
// EXTENSION scene (id 256): the commented-out "Turner Whitted" scene of scenes.rs:113-130 — textured
// Square floor (only if texture 0 is loaded, like `if let Some(t) = textures.get(&0)`), a refracting and a reflecting
// sphere on a sky-blue background — lit by an emissive quad instead of the removed point light.
static inline std::vector<Shape> whitted_shapes(const Texture* tex0) {
  std::vector<Shape> shapes;
  if (tex0) shapes.push_back(Shape::square(Vec3(0.0f, -1.0f, 4.0f), 8.0f, Material::diffuse_texture(tex0)));
  shapes.push_back(Shape::sphere(Vec3(-1.3f, 1.0f, -0.2f), 0.7f, Material::refract(Vec3(0.5f, 1.0f, 0.5f), 1.02f)));
  shapes.push_back(Shape::sphere(Vec3(-0.4f, 0.0f, 1.0f), 0.6f, Material::reflect(Color3(1.0f, 1.0f, 1.0f), 0.3f)));
  Vec3 lc1(-1.0f, 6.0f, -3.0f), lc2(1.0f, 6.0f, -3.0f), lc3(1.0f, 6.0f, -1.0f), lc4(-1.0f, 6.0f, -1.0f);
  shapes.push_back(Shape::triangle(lc3, lc2, lc1, Material::emit(Vec3(16.0f, 16.0f, 16.0f))));
  shapes.push_back(Shape::triangle(lc4, lc3, lc1, Material::emit(Vec3(16.0f, 16.0f, 16.0f))));
  return shapes;
}

static inline std::vector<Shape> mesh_to_triangles(const float* verts, size_t num_vertices) {
  std::vector<Shape> tris;
  size_t nt = num_vertices / 3;
  Material mat = Material::diffuse(Color3(1.0f, 0.4f, 0.4f));
  for (size_t i = 0; i < nt; i++) {
    const float* p = verts + i * 9;
    Vec3 a = Vec3(p[0], p[1], p[2]) * 0.5f, b = Vec3(p[3], p[4], p[5]) * 0.5f, c = Vec3(p[6], p[7], p[8]) * 0.5f;
    Vec3 tr(0.0f, 0.0f, 5.0f);
    tris.push_back(Shape::triangle(a + tr, b + tr, c + tr, mat));
  }
  return tris;
}
// obj_parser.ts:3-51 (+ index.ts:216-220 scale by (8,8,-8) when `client_scale`).
// JS semantics kept: lines split on '\n', fields on single spaces, parseFloat/parseInt
// prefixes; only `v` and `f` matter; non-triangular faces throw.
static inline std::vector<float> parse_obj(const std::string& text, bool client_scale) {
  std::vector<double> vertices;
  std::vector<long> faces;
  size_t pos = 0;
  while (pos <= text.size()) {
    size_t e = text.find('\n', pos);
    if (e == std::string::npos) e = text.size();
    std::string line = text.substr(pos, e - pos);
    pos = e + 1;
    std::vector<std::string> segs;
    size_t s = 0;
    for (;;) { size_t k = line.find(' ', s); if (k == std::string::npos) { segs.push_back(line.substr(s)); break; } segs.push_back(line.substr(s, k - s)); s = k + 1; }
    if (segs[0] == "v") {
      for (int k = 1; k <= 3; k++) vertices.push_back(k < (int)segs.size() ? std::strtod(segs[k].c_str(), nullptr) : NAN);
    } else if (segs[0] == "f") {
      if (segs.size() != 4) throw std::runtime_error("Non-triangular face in OBJ file");
      for (int k = 1; k <= 3; k++) faces.push_back(std::strtol(segs[k].c_str(), nullptr, 10) - 1);
    }
  }
  std::vector<float> out(faces.size() * 3);
  for (size_t i = 0; i < faces.size(); i++)
    for (int c = 0; c < 3; c++) {
      long vi = faces[i] * 3 + c;
      out[i * 3 + c] = (vi >= 0 && (size_t)vi < vertices.size()) ? (float)vertices[vi] : NAN;   // Float32Array store
    }
  if (client_scale) for (size_t i = 0; i < out.size() / 3; i++) { out[i * 3] *= 8.0f; out[i * 3 + 1] *= 8.0f; out[i * 3 + 2] *= -8.0f; }
  return out;
}

}  // namespace ref
