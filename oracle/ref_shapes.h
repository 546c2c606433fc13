// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED. See ref_core.h.
// Restates src/graphics/primitives/{triangle,plane,torus,aa_rect}.rs,
// src/graphics/material.rs and the `roots` 0.0.4 quartic solver (Cargo.lock:81-84,
// sources absent from the reference mount: restated from the crate's published
// algorithm — discriminant test, depressed quartic, Ferrari resolvent cubic).
#pragma once
#include <cstring>
#include "ref_core.h"

namespace ref {

// ---------------------------------------------------------------- Material
// material.rs:16-20 — only Diffuse and Emissive exist at this commit (finding F5)
// EXTENSION (parity unpinned): Reflect / Refract / textured diffuse are named by the
// task but were removed from the reference (only the commented-out Whitted scene, scenes.rs:113-130,
// and the helpers Vec3::reflect / Vec3::exp / Hit::is_entering are left). Their semantics here are ours.
struct Texture {   // texture.rs:9-31
  std::vector<uint8_t> data; uint32_t width = 0, height = 0;
  static uint32_t as_u32(float v) { if (!(v > 0.0f)) return 0u; if (v >= 4294967296.0f) return 0xFFFFFFFFu; return (uint32_t)v; }   // Rust `as u32` saturates
  Color3 at(float u, float v) const {
    uint32_t ix = ((as_u32(std::floor(u * (float)width)) % width) + width) % width;
    uint32_t iy = ((as_u32(std::floor(v * (float)height)) % height) + height) % height;
    const uint8_t* t = &data[(size_t)(iy * width + ix) * 3];
    return Color3((float)t[0] / 255.0f, (float)t[1] / 255.0f, (float)t[2] / 255.0f);
  }
};
enum MatKind : int { MAT_DIFFUSE = 0, MAT_EMISSIVE = 1, MAT_REFLECT = 2, MAT_REFRACT = 3, MAT_DIFFUSE_TEX = 4 };
struct Material {
  bool emissive;
  Color3 color;      // Diffuse / Reflect
  Vec3 intensity;    // Emissive; Refract: absorption
  int kind = MAT_DIFFUSE;
  float param = 0.0f;               // Reflect: reflection share; Refract: index of refraction
  const Texture* tex = nullptr;     // textured diffuse
  static Material diffuse(Color3 c) { Material m; m.emissive = false; m.color = c; m.kind = MAT_DIFFUSE; return m; }
  static Material emit(Vec3 i) { Material m; m.emissive = true; m.intensity = i; m.kind = MAT_EMISSIVE; return m; }
  static Material reflect(Color3 c, float reflection) { Material m = diffuse(c); m.kind = MAT_REFLECT; m.param = reflection; return m; }
  static Material refract(Vec3 absorption, float ior) { Material m = diffuse(Color3(1, 1, 1)); m.kind = MAT_REFRACT; m.intensity = absorption; m.param = ior; return m; }
  static Material diffuse_texture(const Texture* t) { Material m = diffuse(Color3(1, 1, 1)); m.kind = MAT_DIFFUSE_TEX; m.tex = t; return m; }
  bool is_diffuse() const { return kind == MAT_DIFFUSE || kind == MAT_DIFFUSE_TEX; }
  Material evaluate_at(float u, float v) const {   // material.rs:63-70 with the removed texture variant restored
    if (kind != MAT_DIFFUSE_TEX) return *this;
    Material m = diffuse(tex->at(u, v)); return m;
  }
};

// ray.rs:46-63 — Hit::new normalises the normal (again)
struct Hit {
  float distance;
  Vec3 normal;
  Material mat;
  bool is_entering;
  Hit() {}
  Hit(float d, Vec3 n, const Material& m, bool e) : distance(d), normal(normalize(n)), mat(m), is_entering(e) {}
};

// ---------------------------------------------------------------- shared f64 transcendentals
#define F64_FN static inline
static inline unsigned long long f64_bits(double v) { unsigned long long u; std::memcpy(&u, &v, 8); return u; }
static inline double f64_from_bits(unsigned long long u) { double v; std::memcpy(&v, &u, 8); return v; }
#define F64_BITS(a) f64_bits(a)
#define F64_FROM_BITS(b) f64_from_bits(b)

// Shared f64 acos / cos / cbrt for the torus' quartic solver (deviation B11): glibc and CUDA round these
// routines differently in the last place, and a last-place difference in f64 occasionally flips the f32 rounding of a hit
// distance (a handful of the 300 000 museum photons). Both sides therefore evaluate the same f64 + - * / sqrt sequence
// (no FMA contraction on either side): Cody-Waite reduction + Taylor kernels, the asin series, Halley iterations.
// Accuracy: <= 2 ulp against libm (tests/test_oracle_kat.py).
F64_FN double shared_cos64(double x) {
  const double kd = std::floor(x * 0.6366197723675814 + 0.5);
  const double r = ((x - kd * 1.5707963109016418) - kd * 1.5893254773528196e-08) - kd * 6.36831716351095e-25;   // pi/2 in three parts, kd * part 1 is exact
  const double z = r * r;
  const int q = (int)kd & 3;
  if (q & 1) {   // +- sin r
    const double s = r + r * z * (-0.16666666666666666 + z * (0.008333333333333333 + z * (-0.0001984126984126984 + z * (2.7557319223985893e-06 + z * (-2.505210838544172e-08 +
                     z * (1.6059043836821613e-10 + z * (-7.647163731819816e-13 + z * (2.8114572543455206e-15 + z * -8.22063524662433e-18))))))));
    return q == 1 ? -s : s;
  }
  const double c = 1.0 + z * (-0.5 + z * (0.041666666666666664 + z * (-0.001388888888888889 + z * (2.48015873015873e-05 + z * (-2.755731922398589e-07 +
                   z * (2.08767569878681e-09 + z * (-1.1470745597729725e-11 + z * (4.779477332387385e-14 + z * -1.5619206968586225e-16))))))));
  return q == 0 ? c : -c;
}
F64_FN double shared_asin_small64(double x) {   // |x| <= 0.5: x + x z (c1 + z (c2 + ...)), 27 terms of the series (next term 1e-19 at 0.5)
  const double z = x * x;
  double p = 0.0019650336162772837;
  const double c[26] = {0.0020776610325181676, 0.0022014739737101384, 0.002338091892111975, 0.0024894486782468836, 0.00265787063820729, 0.002846178401108942,
                        0.0030578216492580306, 0.003297059503473485, 0.0035692053938259347, 0.003880964558837669, 0.004240907093679363, 0.004660143486915096,
                        0.005153309682319905, 0.005740037670841924, 0.006447210311889649, 0.0073125258735988454, 0.008390335809616815, 0.009761609529194078,
                        0.011551800896139705, 0.01396484375, 0.017352764423076924, 0.022372159090909092, 0.030381944444444444, 0.044642857142857144, 0.075,
                        0.16666666666666666};
  for (int i = 0; i < 26; i++) p = c[i] + z * p;
  return x + x * z * p;
}
F64_FN double shared_acos64(double x) {
  if (x > 0.5) return 2.0 * shared_asin_small64(std::sqrt((1.0 - x) * 0.5));
  if (x < -0.5) return 3.141592653589793 - 2.0 * shared_asin_small64(std::sqrt((1.0 + x) * 0.5));
  return 1.5707963267948966 - shared_asin_small64(x);
}
F64_FN double shared_cbrt64(double x) {
  if (x == 0.0 || x != x) return x;
  double a = x < 0.0 ? -x : x;
  if (a > 1.7976931348623157e308) return x;
  double scale = 1.0;
  if (a < 2.2250738585072014e-308) { a *= 18014398509481984.0; scale = 3.814697265625e-06; }   // subnormal: 2^54, 2^-18
  double t = F64_FROM_BITS(F64_BITS(a) / 3ull + 0x2A9F7893782DA1CEull);                          // 3 % initial guess
  for (int i = 0; i < 3; i++) { const double t3 = t * t * t; t = t * ((t3 + a + a) / (t3 + t3 + a)); }   // Halley, cubic convergence: 3e-2 -> 3e-5 -> 2e-14 -> 0
  t = t - (t * t * t - a) / (3.0 * t * t);
  return (x < 0.0 ? -t : t) * scale;
}
#undef F64_FN
#undef F64_BITS
#undef F64_FROM_BITS

// ---------------------------------------------------------------- roots 0.0.4 (restated)
struct Roots {
  int n = 0;
  double v[4];
  // roots: Roots::add_new_root keeps the set sorted ascending and drops exact duplicates
  void add(double x) {
    for (int i = 0; i < n; i++) if (v[i] == x) return;
    if (n == 4) return;
    int i = n;
    while (i > 0 && v[i - 1] > x) { v[i] = v[i - 1]; i--; }
    v[i] = x;
    n++;
  }
};
static inline Roots roots_quadratic_normalized(double a1, double a0) {
  Roots r;
  double disc = a1 * a1 - 4.0 * a0;
  if (disc < 0.0) return r;
  double a1_div_2 = a1 / 2.0;
  if (disc == 0.0) { r.add(-a1_div_2); return r; }
  double sq = std::sqrt(disc);
  r.add(-a1_div_2 - sq / 2.0);
  r.add(-a1_div_2 + sq / 2.0);
  return r;
}
static inline Roots roots_quadratic(double a2, double a1, double a0) {
  Roots r;
  if (a2 == 0.0) { if (a1 != 0.0) r.add(-a0 / a1); return r; }
  double disc = a1 * a1 - 4.0 * a2 * a0;
  if (disc < 0.0) return r;
  double a2x2 = 2.0 * a2;
  if (disc == 0.0) { r.add(-a1 / a2x2); return r; }
  double sq = std::sqrt(disc);
  r.add((-a1 - sq) / a2x2);
  r.add((-a1 + sq) / a2x2);
  return r;
}
static inline Roots roots_cubic_normalized(double a2, double a1, double a0) {
  Roots out;
  double q = (3.0 * a1 - a2 * a2) / 9.0;
  double r = (9.0 * a2 * a1 - 27.0 * a0 - 2.0 * a2 * a2 * a2) / 54.0;
  double q3 = q * q * q;
  double d = q3 + r * r;
  double a2_div_3 = a2 / 3.0;
  if (d < 0.0) {
    double phi_3 = shared_acos64(r / std::sqrt(-q3)) / 3.0;
    double sqrt_q_2 = 2.0 * std::sqrt(-q);
    const double two_third_pi = 2.0943951023931954923;
    out.add(sqrt_q_2 * shared_cos64(phi_3) - a2_div_3);
    out.add(sqrt_q_2 * shared_cos64(phi_3 - two_third_pi) - a2_div_3);
    out.add(sqrt_q_2 * shared_cos64(phi_3 + two_third_pi) - a2_div_3);
  } else {
    double sqrt_d = std::sqrt(d);
    double s = shared_cbrt64(r + sqrt_d);
    double t = shared_cbrt64(r - sqrt_d);
    out.add(s + t - a2_div_3);
    if (s == t && s + t != 0.0) out.add(-(s + t) / 2.0 - a2_div_3);
  }
  return out;
}
static inline Roots roots_cubic(double a3, double a2, double a1, double a0) {
  if (a3 == 0.0) return roots_quadratic(a2, a1, a0);
  if (a2 == 0.0 && a1 == 0.0 && a0 == 0.0) { Roots r; r.add(0.0); return r; }
  return roots_cubic_normalized(a2 / a3, a1 / a3, a0 / a3);
}
static inline Roots roots_biquadratic(double a4, double a2, double a0) {
  Roots out;
  Roots q = roots_quadratic(a4, a2, a0);
  for (int i = 0; i < q.n; i++) {
    double x = q.v[i];
    if (x > 0.0) { double s = std::sqrt(x); out.add(-s); out.add(s); }
    else if (x == 0.0) out.add(0.0);
  }
  return out;
}
static inline Roots roots_quartic_depressed(double a2, double a1, double a0) {
  if (a1 == 0.0) return roots_biquadratic(1.0, a2, a0);
  if (a0 == 0.0) { Roots r = roots_cubic_normalized(0.0, a2, a1); r.add(0.0); return r; }
  double a2_pow_2 = a2 * a2;
  double a1_div_2 = a1 / 2.0;
  double b2 = a2 * 5.0 / 2.0;
  double b1 = 2.0 * a2_pow_2 - a0;
  double b0 = (a2_pow_2 * a2 - a2 * a0 - a1_div_2 * a1_div_2) / 2.0;
  Roots res = roots_cubic_normalized(b2, b1, b0);
  double y = res.v[res.n - 1];          // the largest resolvent root
  double a2_plus_2y = a2 + 2.0 * y;
  Roots out;
  if (a2_plus_2y > 0.0) {
    double s = std::sqrt(a2_plus_2y);
    double q0a = a2 + y - a1_div_2 / s;
    double q0b = a2 + y + a1_div_2 / s;
    Roots ra = roots_quadratic_normalized(s, q0a);
    Roots rb = roots_quadratic_normalized(-s, q0b);
    for (int i = 0; i < ra.n; i++) out.add(ra.v[i]);
    for (int i = 0; i < rb.n; i++) out.add(rb.v[i]);
  }
  return out;
}
static inline Roots roots_quartic(double a4, double a3, double a2, double a1, double a0) {
  if (a4 == 0.0) return roots_cubic(a3, a2, a1, a0);
  if (a0 == 0.0) { Roots r = roots_cubic(a4, a3, a2, a1); r.add(0.0); return r; }
  if (a1 == 0.0 && a3 == 0.0) return roots_biquadratic(a4, a2, a0);
  double discriminant =
      a4 * a0 * a4 * (256.0 * a4 * a0 * a0 + a1 * (144.0 * a2 * a1 - 192.0 * a3 * a0)) +
      a4 * a0 * a2 * a2 * (16.0 * a2 * a2 - 80.0 * a3 * a1 - 128.0 * a4 * a0) +
      (a3 * a3 * (a4 * a0 * (144.0 * a2 * a0 - 6.0 * a1 * a1) +
                  (a0 * (18.0 * a3 * a2 * a1 - 27.0 * a3 * a3 * a0 - 4.0 * a2 * a2 * a2) +
                   a1 * a1 * (a2 * a2 - 4.0 * a3 * a1)))) +
      a4 * a1 * a1 * (18.0 * a3 * a2 * a1 - 27.0 * a4 * a1 * a1 - 4.0 * a2 * a2 * a2);
  double pp = 8.0 * a4 * a2 - 3.0 * a3 * a3;
  double rr = a3 * a3 * a3 + 8.0 * a4 * a4 * a1 - 4.0 * a4 * a3 * a2;
  double delta0 = a2 * a2 - 3.0 * a3 * a1 + 12.0 * a4 * a0;
  double dd = 64.0 * a4 * a4 * a4 * a0 - 16.0 * a4 * a4 * a2 * a2 + 16.0 * a4 * a3 * a3 * a2 -
              16.0 * a4 * a4 * a3 * a1 - 3.0 * a3 * a3 * a3 * a3;
  Roots out;
  if (discriminant == 0.0) {
    bool triple = delta0 == 0.0;
    bool quadruple = triple && dd == 0.0;
    bool no_roots = dd == 0.0 && pp > 0.0 && rr == 0.0;
    if (quadruple) { out.add(-a3 / (4.0 * a4)); return out; }
    if (triple) {
      double x0 = (-72.0 * a4 * a4 * a0 + 10.0 * a4 * a2 * a2 - 3.0 * a3 * a3 * a2) /
                  (9.0 * (8.0 * a4 * a4 * a1 - 4.0 * a4 * a3 * a2 + a3 * a3 * a3));
      out.add(x0);
      out.add(-(a3 / a4 + 3.0 * x0));
      return out;
    }
    if (no_roots) return out;
  } else {
    if (discriminant > 0.0 && (pp > 0.0 || dd > 0.0)) return out;
  }
  double a4_pow_2 = a4 * a4, a4_pow_3 = a4_pow_2 * a4, a4_pow_4 = a4_pow_2 * a4_pow_2;
  double p = pp / (8.0 * a4_pow_2);
  double q = rr / (8.0 * a4_pow_3);
  double r = (dd + 16.0 * a4_pow_2 * (12.0 * a0 * a4 - 3.0 * a1 * a3 + a2 * a2)) / (256.0 * a4_pow_4);
  Roots dep = roots_quartic_depressed(p, q, r);
  for (int i = 0; i < dep.n; i++) out.add(dep.v[i] - a3 / (4.0 * a4));
  return out;
}

// ---------------------------------------------------------------- Shape
// The reference uses Rc<dyn Tracable>; a tagged struct is the same thing flattened.
enum ShapeType : int { SH_TRIANGLE = 0, SH_PLANE = 1, SH_TORUS = 2, SH_AARECT = 3, SH_SPHERE = 4, SH_SQUARE = 5 };

struct Shape {
  ShapeType type;
  Material mat;
  Vec3 v0, v1, v2;          // triangle
  Vec3 location, normal;    // plane (location, normal) / torus (location)
  float big_r, small_r;     // torus; sphere: big_r = radius; square: big_r = size
  float x_min, x_max, y_min, y_max, z_min, z_max;   // aa_rect (note: min/max interleaved, aa_rect.rs:8-16)
  int source_index = -1;    // position in the scene's original shape list (debug / tests)

  static Shape triangle(Vec3 a, Vec3 b, Vec3 c, Material m) {
    Shape s{}; s.type = SH_TRIANGLE; s.v0 = a; s.v1 = b; s.v2 = c; s.mat = m; return s;
  }
  static Shape plane(Vec3 loc, Vec3 n, Material m) {
    Shape s{}; s.type = SH_PLANE; s.location = loc; s.normal = n; s.mat = m; return s;
  }
  static Shape torus(Vec3 loc, float R, float r, Material m) {
    Shape s{}; s.type = SH_TORUS; s.location = loc; s.big_r = R; s.small_r = r; s.mat = m; return s;
  }
  static Shape aarect(float x0, float x1, float y0, float y1, float z0, float z1, Material m) {
    Shape s{}; s.type = SH_AARECT; s.x_min = x0; s.x_max = x1; s.y_min = y0; s.y_max = y1; s.z_min = z0; s.z_max = z1; s.mat = m; return s;
  }

  static Shape sphere(Vec3 loc, float radius, Material m) {   // sphere.rs:17-22
    Shape s{}; s.type = SH_SPHERE; s.location = loc; s.big_r = radius; s.mat = m; return s;
  }
  static Shape square(Vec3 loc, float size, Material m) {     // square.rs:17-22
    Shape s{}; s.type = SH_SQUARE; s.location = loc; s.big_r = size; s.mat = m; return s;
  }

  bool is_emissive() const { return mat.emissive; }

  // Bounded::aabb — false for infinite shapes (plane.rs:33-36)
  bool aabb(AABB* out) const {
    switch (type) {
      case SH_TRIANGLE: {   // triangle.rs:48-66
        float xmn = fmin_(fmin_(v0.x, v1.x), v2.x), ymn = fmin_(fmin_(v0.y, v1.y), v2.y), zmn = fmin_(fmin_(v0.z, v1.z), v2.z);
        float xmx = fmax_(fmax_(v0.x, v1.x), v2.x), ymx = fmax_(fmax_(v0.y, v1.y), v2.y), zmx = fmax_(fmax_(v0.z, v1.z), v2.z);
        float pad = 0.1f * EPSILON;
        *out = AABB(xmn - pad, ymn - pad, zmn - pad, xmx + pad, ymx + pad, zmx + pad);
        return true;
      }
      case SH_TORUS: {      // torus.rs:33-52
        float r = big_r + small_r;
        *out = AABB(location.x - r, location.y - small_r, location.z - r, location.x + r, location.y + small_r, location.z + r);
        return true;
      }
      case SH_AARECT:       // aa_rect.rs:57-67
        *out = AABB(x_min, y_min, z_min, x_max, y_max, z_max);
        return true;
      case SH_SPHERE: {     // sphere.rs:31-36
        float r = big_r;
        *out = AABB(location.x - r, location.y - r, location.z - r, location.x + r, location.y + r, location.z + r);
        return true;
      }
      case SH_SQUARE: {     // square.rs:31-44 (flat box)
        float hs = big_r * 0.5f;
        *out = AABB(location.x - hs, location.y, location.z - hs, location.x + hs, location.y, location.z + hs);
        return true;
      }
      default: return false;
    }
  }
  // Bounded::location (ray.rs:77-83 default = AABB centre; torus.rs:28-30; aa_rect.rs:48-54)
  bool centroid(Vec3* out) const {
    switch (type) {
      case SH_TRIANGLE: { AABB b; aabb(&b); *out = b.center(); return true; }
      case SH_TORUS: case SH_SPHERE: case SH_SQUARE: *out = location; return true;   // torus.rs:28-30, sphere.rs:26-28, square.rs:26-28
      case SH_AARECT: *out = Vec3(0.5f * (x_min + x_max), 0.5f * (y_min + y_max), 0.5f * (z_min + z_max)); return true;
      default: return false;
    }
  }

  // triangle.rs:70-87 (Heron, evaluated on every call in the reference)
  float surface_area() const {
    if (type != SH_TRIANGLE) throw std::runtime_error("Not implemented");
    float a = dis(v0, v1), b = dis(v1, v2), c = dis(v2, v0);
    float s = (a + b + c) * 0.5f;
    return std::sqrt(s * (s - a) * (s - b) * (s - c));
  }
  // triangle.rs:91-114 — returns (point, normal, intensity); quirk q2: random normal flip
  void pick_random(Rng& rng, Vec3* p, Vec3* n_out, Vec3* intensity) const {
    if (type != SH_TRIANGLE) throw std::runtime_error("Not implemented");
    float r1 = rng.next();
    float r2 = rng.next();
    float r1_sqrt = std::sqrt(r1);
    Vec3 p_hit = (1.0f - r1_sqrt) * v0 + (r1_sqrt * (1.0f - r2)) * v1 + (r2 * r1_sqrt) * v2;
    Vec3 n = normalize(cross(v1 - v0, v2 - v0));
    if (rng.next() > 0.5f) n = -n;
    if (mat.emissive) { *p = p_hit; *n_out = n; *intensity = mat.intensity; }
    else { *p = Vec3(); *n_out = Vec3(); *intensity = Vec3(); }
  }

  static bool is_approx_left_of(Vec3 a, Vec3 b, Vec3 n, Vec3 p) {   // triangle.rs:41-45
    Vec3 edge = b - a;
    Vec3 ap = p - a;
    return dot(n, cross(edge, ap)) + 0.1f * EPSILON >= 0.0f;
  }

  // Tracable::trace_simple
  bool trace_simple(const Ray& ray, float* t_out) const {
    switch (type) {
      case SH_TRIANGLE: {   // triangle.rs:159-191
        Vec3 n = cross(v1 - v0, v2 - v0);
        float n_dot_d = dot(n, ray.dir);
        if (n_dot_d == 0.0f) return false;
        float orig_dis = dot(n, v0);
        float t = (orig_dis - dot(n, ray.origin)) / n_dot_d;
        if (t <= 0.0f) return false;
        n = normalize(n);
        Vec3 p = ray.at(t);
        if (is_approx_left_of(v0, v1, n, p) && is_approx_left_of(v1, v2, n, p) && is_approx_left_of(v2, v0, n, p)) { *t_out = t; return true; }
        return false;
      }
      case SH_PLANE: {      // plane.rs:80-99
        float n_dot_dir = dot(normal, ray.dir);
        if (n_dot_dir == 0.0f) return false;
        float o_distance = dot(normal, location);
        float t = (o_distance - dot(normal, ray.origin)) / n_dot_dir;
        if (t <= 0.0f) return false;
        *t_out = t;
        return true;
      }
      case SH_AARECT: {     // aa_rect.rs:142-174 (own 1/dir, strict tmin >= tmax, strict > 0)
        float invdx = 1.0f / ray.dir.x, invdy = 1.0f / ray.dir.y, invdz = 1.0f / ray.dir.z;
        float tx1 = (x_min - ray.origin.x) * invdx, tx2 = (x_max - ray.origin.x) * invdx;
        float ty1 = (y_min - ray.origin.y) * invdy, ty2 = (y_max - ray.origin.y) * invdy;
        float tz1 = (z_min - ray.origin.z) * invdz, tz2 = (z_max - ray.origin.z) * invdz;
        float tmin = fmax_(fmax_(fmin_(tx1, tx2), fmin_(ty1, ty2)), fmin_(tz1, tz2));
        float tmax = fmin_(fmin_(fmax_(tx1, tx2), fmax_(ty1, ty2)), fmax_(tz1, tz2));
        if (tmin >= tmax) return false;
        if (tmin > 0.0f) { *t_out = tmin; return true; }
        if (tmax > 0.0f) { *t_out = tmax; return true; }
        return false;
      }
      case SH_TORUS: case SH_SQUARE: {   // ray.rs:110-116 default: trace().distance
        Hit h;
        if (!trace(ray, &h)) return false;
        *t_out = h.distance;
        return true;
      }
      case SH_SPHERE: {     // sphere.rs:104-131
        float t;
        if (!sphere_t(ray, &t, nullptr)) return false;
        *t_out = t;
        return true;
      }
    }
    return false;
  }
  // sphere.rs:55-82 / 106-128: algebraic solution with a = 1
  bool sphere_t(const Ray& ray, float* t_out, bool* entering) const {
    float a = 1.0f;
    float b = 2.0f * dot(ray.dir, ray.origin - location);
    float c = dot(ray.origin - location, ray.origin - location) - big_r * big_r;
    float d = b * b - 4.0f * a * c;
    if (d < 0.0f) return false;
    float d_sqrt = std::sqrt(d);
    float t0 = (-b + d_sqrt) / (2.0f * a);
    float t1 = (-b - d_sqrt) / (2.0f * a);
    float t = fmin_(t0, t1);
    bool ent = true;
    if (t <= 0.0f) {
      t = fmax_(t0, t1);
      if (t <= 0.0f) return false;
      ent = false;
    }
    *t_out = t;
    if (entering) *entering = ent;
    return true;
  }

  // Tracable::trace
  bool trace(const Ray& ray, Hit* out) const {
    switch (type) {
      case SH_TRIANGLE: {   // triangle.rs:116-157
        Vec3 n = cross(v1 - v0, v2 - v0);
        float n_dot_d = dot(n, ray.dir);
        if (n_dot_d == 0.0f) return false;
        float orig_dis = dot(n, v0);
        float t = (orig_dis - dot(n, ray.origin)) / n_dot_d;
        if (t <= 0.0f) return false;
        n = normalize(n);
        Vec3 p = ray.at(t);
        if (is_approx_left_of(v0, v1, n, p) && is_approx_left_of(v1, v2, n, p) && is_approx_left_of(v2, v0, n, p)) {
          if (n_dot_d > 0.0f) *out = Hit(t, -n, mat, false);
          else *out = Hit(t, n, mat, true);
          return true;
        }
        return false;
      }
      case SH_PLANE: {      // plane.rs:45-77
        Vec3 nrm = normal;
        float n_dot_dir = dot(nrm, ray.dir);
        if (n_dot_dir == 0.0f) return false;
        float o_distance = dot(nrm, location);
        float t = (o_distance - dot(nrm, ray.origin)) / n_dot_dir;
        if (t <= 0.0f) return false;
        if (n_dot_dir > 0.0f) nrm = -nrm;
        *out = Hit(t, nrm, mat, true);
        return true;
      }
      case SH_AARECT: {     // aa_rect.rs:71-139 — face normal by float equality
        float invdx = 1.0f / ray.dir.x, invdy = 1.0f / ray.dir.y, invdz = 1.0f / ray.dir.z;
        float tx1 = (x_min - ray.origin.x) * invdx, tx2 = (x_max - ray.origin.x) * invdx;
        float ty1 = (y_min - ray.origin.y) * invdy, ty2 = (y_max - ray.origin.y) * invdy;
        float tz1 = (z_min - ray.origin.z) * invdz, tz2 = (z_max - ray.origin.z) * invdz;
        float tmin = fmax_(fmax_(fmin_(tx1, tx2), fmin_(ty1, ty2)), fmin_(tz1, tz2));
        float tmax = fmin_(fmin_(fmax_(tx1, tx2), fmax_(ty1, ty2)), fmax_(tz1, tz2));
        if (tmin >= tmax) return false;
        if (tmin > 0.0f) {
          Vec3 nn;
          if (tmin == tx1) nn = Vec3(-1, 0, 0);
          else if (tmin == tx2) nn = Vec3(1, 0, 0);
          else if (tmin == ty1) nn = Vec3(0, -1, 0);
          else if (tmin == ty2) nn = Vec3(0, 1, 0);
          else if (tmin == tz1) nn = Vec3(0, 0, -1);
          else nn = Vec3(0, 0, 1);
          *out = Hit(tmin, nn, mat, true);
          return true;
        }
        if (tmax > 0.0f) {
          Vec3 nn;
          if (tmax == tx1) nn = Vec3(1, 0, 0);
          else if (tmax == tx2) nn = Vec3(-1, 0, 0);
          else if (tmax == ty1) nn = Vec3(0, 1, 0);
          else if (tmax == ty2) nn = Vec3(0, -1, 0);
          else if (tmax == tz1) nn = Vec3(0, 0, 1);
          else nn = Vec3(0, 0, -1);
          *out = Hit(tmax, nn, mat, false);
          return true;
        }
        return false;
      }
      case SH_TORUS: {      // torus.rs:61-126 — f64 quartic
        double a = (double)big_r, b = (double)small_r;
        Vec3 d = ray.origin - location;
        Vec3 e = ray.dir;
        double dx = d.x, dy = d.y, dz = d.z, ex = e.x, ey = e.y, ez = e.z;
        double g = 4.0 * a * a * (ex * ex + ez * ez);
        double h = 8.0 * a * a * (dx * ex + dz * ez);
        double i = 4.0 * a * a * (dx * dx + dz * dz);
        double j = ex * ex + ey * ey + ez * ez;
        double k = 2.0 * (dx * ex + dy * ey + dz * ez);
        double l = dx * dx + dy * dy + dz * dz + a * a - b * b;
        Roots rs = roots_quartic(j * j, 2.0 * j * k, 2.0 * j * l + k * k - g, 2.0 * k * l - h, l * l - i);
        double pos[4]; int np = 0;
        for (int q = 0; q < rs.n; q++) if (rs.v[q] >= 0.0001) pos[np++] = rs.v[q];   // torus.rs:130-139
        if (np == 0) return false;
        double closest = pos[0];
        for (int q = 1; q < np; q++) closest = std::fmin(closest, pos[q]);
        double px = (double)d.x + (double)e.x * closest;
        double py = (double)d.y + (double)e.y * closest;
        double pz = (double)d.z + (double)e.z * closest;
        double alpha = 1.0 - a / std::sqrt(px * px + pz * pz);
        Vec3 n = unit((float)(alpha * px), (float)py, (float)(alpha * pz));
        if (np % 2 == 1) *out = Hit((float)closest, -n, mat, false);
        else *out = Hit((float)closest, n, mat, true);
        return true;
      }
      case SH_SPHERE: {     // sphere.rs:49-101 (textured spheres are not supported: atan2 / asin are libm-dependent)
        float t; bool ent;
        if (!sphere_t(ray, &t, &ent)) return false;
        Vec3 nrm = (ray.at(t) - location) / big_r;
        if (mat.kind == MAT_DIFFUSE_TEX) throw std::runtime_error("textured sphere not supported");
        *out = Hit(t, ent ? nrm : -nrm, mat, ent);
        return true;
      }
      case SH_SQUARE: {     // square.rs:56-99
        float n_dot_dir = ray.dir.y;
        if (n_dot_dir == 0.0f) return false;
        float t = (location.y - ray.origin.y) / n_dot_dir;
        if (t <= 0.0f) return false;
        Vec3 hit = ray.at(t);
        float dx = std::fabs(hit.x - location.x), dz = std::fabs(hit.z - location.z);
        if (2.0f * dx >= big_r || 2.0f * dz >= big_r) return false;
        Vec3 nrm = n_dot_dir > 0.0f ? Vec3(0.0f, -1.0f, 0.0f) : Vec3(0.0f, 1.0f, 0.0f);
        float u = (hit.x - location.x) / big_r + 0.5f, v = (hit.z - location.z) / big_r + 0.5f;
        *out = Hit(t, nrm, mat.evaluate_at(u, v), true);
        return true;
      }
    }
    return false;
  }
};

}  // namespace ref
