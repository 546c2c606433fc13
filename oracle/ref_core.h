// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED.
//
// CPU restatement of sourcedennis/wasm-pathtracer's hot path (citations are
// relative to the reference checkout, `src/...`). The reference ships no tests,
// no golden vectors and cannot be compiled in this environment (no Rust
// toolchain), so nothing here is pinned against reference OUTPUT; the only
// reference artefact that corroborates it is the museum light-colour order in
// banner.png (tests/test_oracle_kat.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may build, link or call anything under oracle/. The product
// library (wasm_pathtracer_b200/csrc) never includes these headers.
//
// Build flags that matter: -ffp-contract=off -fno-fast-math (Rust never
// contracts a*b+c into an FMA and uses IEEE div/sqrt).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <string>
#include <stdexcept>
#include <algorithm>

namespace ref {

// src/math/mod.rs:11
static const float EPSILON = 0.0002f;
static const float PI_F = 3.14159265358979323846f;   // std::f32::consts::PI
static const float INF_F = std::numeric_limits<float>::infinity();

// Rust f32::min / f32::max return the non-NaN operand; fminf/fmaxf do the same.
static inline float fmin_(float a, float b) { return std::fmin(a, b); }
static inline float fmax_(float a, float b) { return std::fmax(a, b); }
// src/math/mod.rs:13-15   max_val.min( min_val.max( x ) )
static inline float clampf(float x, float lo, float hi) { return fmin_(hi, fmax_(lo, x)); }

// ---------------------------------------------------------------- Vec3
// src/math/vec3.rs
struct Vec3 {
  float x, y, z;
  Vec3() : x(0), y(0), z(0) {}
  Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
static inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
static inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
// vec3.rs:151-165: both Vec3*f32 and f32*Vec3 compute `multiplier * component`
static inline Vec3 operator*(Vec3 a, float m) { return Vec3(m * a.x, m * a.y, m * a.z); }
static inline Vec3 operator*(float m, Vec3 a) { return Vec3(m * a.x, m * a.y, m * a.z); }
static inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline Vec3 operator/(Vec3 a, float d) { return Vec3(a.x / d, a.y / d, a.z / d); }
static inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // vec3.rs:32-34
static inline Vec3 cross(Vec3 a, Vec3 t) {                                               // vec3.rs:57-62
  return Vec3(a.y * t.z - a.z * t.y, a.z * t.x - a.x * t.z, a.x * t.y - a.y * t.x);
}
static inline float len_sq(Vec3 a) { return dot(a, a); }
static inline float len(Vec3 a) { return std::sqrt(len_sq(a)); }
static inline Vec3 normalize(Vec3 a) { return a * (1.0f / len(a)); }                     // vec3.rs:27-29
static inline Vec3 unit(float x, float y, float z) { return normalize(Vec3(x, y, z)); }
static inline float dis_sq(Vec3 a, Vec3 b) { return len_sq(a - b); }
static inline float dis(Vec3 a, Vec3 b) { return len(a - b); }
// vec3.rs:37-54
static inline Vec3 orthogonal(Vec3 s) {
  if (std::fabs(s.z) > 0.1f) {
    float v1 = 1.0f, v2 = 1.0f;
    float v3 = -(s.x * v1 + s.y * v2) / s.z;
    return unit(v1, v2, v3);
  } else if (std::fabs(s.x) > 0.1f) {
    float v2 = 1.0f, v3 = 1.0f;
    float v1 = -(s.y * v2 + s.z * v3) / s.x;
    return unit(v1, v2, v3);
  } else {
    float v1 = 1.0f, v3 = 1.0f;
    float v2 = -(s.x * v1 + s.z * v3) / s.y;
    return unit(v1, v2, v3);
  }
}
// vec3.rs:95-119. sin/cos of the camera angles use libm (evaluated per sample in
// the reference; the value is the same every time).
static inline Vec3 rot_y(Vec3 v, float angle) {
  float c = std::cos(angle), s = std::sin(angle);
  return Vec3(c * v.x + s * v.z, v.y, -s * v.x + c * v.z);
}
static inline Vec3 rot_x(Vec3 v, float angle) {
  float c = std::cos(angle), s = std::sin(angle);
  return Vec3(v.x, c * v.y - s * v.z, s * v.y + c * v.z);
}

// ---------------------------------------------------------------- Color3
// src/graphics/color3.rs — every constructor clamps to [0,1]
struct Color3 {
  float red, green, blue;
  Color3() : red(0), green(0), blue(0) {}
  Color3(float r, float g, float b)
      : red(clampf(r, 0.0f, 1.0f)), green(clampf(g, 0.0f, 1.0f)), blue(clampf(b, 0.0f, 1.0f)) {}
  Vec3 to_vec3() const { return Vec3(red, green, blue); }
};
static inline Color3 operator*(Color3 c, float m) { return Color3(m * c.red, m * c.green, m * c.blue); }
static inline Color3 operator/(Color3 c, float v) { return c * (1.0f / v); }   // color3.rs:89-95

// ---------------------------------------------------------------- Rng
// src/rng.rs
struct Rng {
  uint32_t state;
  Rng() : state(0xBABABEBEu) {}
  explicit Rng(uint32_t s) : state(s) {}
  uint32_t next_u32() {                       // rng.rs:40-47
    uint32_t x = state;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    state = x;
    return x;
  }
  // rng.rs:19-21 — `0xFFFFFFFFu32 as f32` is 2^32, so the scale is exactly 2^-32
  float next() { return (float)next_u32() * (1.0f / 4294967296.0f); }
  // rng.rs:25-38 — quirk q3: single-element range returns 0 (not `low`), no draw
  size_t next_in_range(size_t low, size_t high) {
    if (high <= low) throw std::runtime_error("Invalid range");
    if (high == low + 1) return 0;
    float f = next();
    if (f == 1.0f) return high - 1;
    return (size_t)std::floor(f * (float)(high - low)) + low;
  }
  // rng.rs:50-68
  Vec3 next_hemisphere(Vec3 normal) {
    float x, y, z;
    for (;;) {
      x = next() * 2.0f - 1.0f;
      y = next() * 2.0f - 1.0f;
      z = next() * 2.0f - 1.0f;
      float ls = x * x + y * y + z * z;
      if (!(ls > 1.0f)) break;
    }
    Vec3 v = unit(x, y, z);
    if (dot(v, normal) < 0.0f) return -v;
    return v;
  }
  template <class T> void shuffle(std::vector<T>& xs) {   // rng.rs:70-75
    for (size_t i = 0; i < xs.size(); i++) {
      size_t j = next_in_range(0, xs.size());
      std::swap(xs[i], xs[j]);
    }
  }
};

// ---------------------------------------------------------------- mode-B stream contract
// Not in the reference (finding F8: one sequential stream cannot be reproduced by a
// parallel renderer). DESIGN.md "RNG contract": every path / photon shot owns an
// xorshift32 stream whose seed is a hash of (index, sample, stream id, base seed).
enum StreamId : uint32_t { STREAM_PATH = 1, STREAM_PHOTON = 2, STREAM_PIXEL = 3 };
static inline uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
static inline uint32_t stream_seed(uint32_t index, uint32_t sample, uint32_t stream, uint32_t base) {
  uint32_t s = mix32(index + mix32(sample + mix32(stream ^ base)));
  return s == 0 ? 0xBABABEBEu : s;
}

// Shared trig for mode B ("trig = shared"): cos/sin of a in [0, 2*pi] built only from
// f32 + - * so that CPU (-ffp-contract=off) and GPU (-fmad=false) agree bit for bit.
// The reference calls f32::cos/sin (material.rs:103-105), whose last-ulp behaviour is
// platform libm; "trig = libm" keeps that.
static inline void shared_sincos(float a, float* s_out, float* c_out) {
  int k = (int)(a * 0.63661977f + 0.5f);
  float fk = (float)k;
  float r = (a - fk * 1.5707963f) - fk * (-4.371139e-8f);
  float r2 = r * r;
  float s = r + r * r2 * (-1.6666654611e-1f + r2 * (8.3321608736e-3f + r2 * (-1.9515295891e-4f)));
  float c = (1.0f - 0.5f * r2) + r2 * r2 * (4.166664568298827e-2f + r2 * (-1.388731625493765e-3f + r2 * 2.443315711809948e-5f));
  switch (k & 3) {
    case 0: *s_out = s;  *c_out = c;  break;
    case 1: *s_out = c;  *c_out = -s; break;
    case 2: *s_out = -s; *c_out = -c; break;
    default: *s_out = -c; *c_out = s; break;
  }
}

// EXTENSION: e^-x for x >= 0 from f32 + - * and exponent bits only, identical on CPU and GPU
// (the reference keeps Vec3::exp, vec3.rs:88-92, for Beer's law but no material uses it any more).
static inline float shared_exp_neg(float x) {
  if (!(x < 87.0f)) return 0.0f;
  float kf = (float)(int)(x * 1.44269504f + 0.5f);            // round(x / ln 2)
  float r = (kf * 0.693145751953125f - x) + kf * 1.42860682030941723212e-6f;   // -(x - k ln2), Cody-Waite split
  float p = 1.0f + r * (1.0f + r * (0.5f + r * (0.16666667f + r * (0.041666668f + r * (0.008333334f + r * 0.0013888889f)))));
  uint32_t bits = (uint32_t)(127 - (int)kf) << 23;             // 2^-k
  float scale; std::memcpy(&scale, &bits, 4);
  return p * scale;
}

// ---------------------------------------------------------------- Ray
// src/graphics/ray.rs:22-39
struct Ray {
  Vec3 origin, dir, inv_dir;
  Ray() {}
  Ray(Vec3 o, Vec3 d) : origin(o), dir(d), inv_dir(1.0f / d.x, 1.0f / d.y, 1.0f / d.z) {}
  Vec3 at(float distance) const { return origin + distance * dir; }
};

// ---------------------------------------------------------------- AABB
// src/graphics/aabb.rs
struct AABB {
  float x_min, y_min, z_min, x_max, y_max, z_max;
  AABB() : x_min(0), y_min(0), z_min(0), x_max(0), y_max(0), z_max(0) {}
  AABB(float a, float b, float c, float d, float e, float f)
      : x_min(a), y_min(b), z_min(c), x_max(d), y_max(e), z_max(f) {}
  float x_size() const { return x_max - x_min; }
  float y_size() const { return y_max - y_min; }
  float z_size() const { return z_max - z_min; }
  float surface() const {                                   // aabb.rs:72-78
    float xs = x_max - x_min, ys = y_max - y_min, zs = z_max - z_min;
    return 2.0f * (xs * ys + xs * zs + ys * zs);
  }
  Vec3 center() const {                                     // aabb.rs:81-87
    return Vec3(0.5f * (x_min + x_max), 0.5f * (y_min + y_max), 0.5f * (z_min + z_max));
  }
  AABB join(const AABB& o) const {                          // aabb.rs:90-100
    return AABB(fmin_(x_min, o.x_min), fmin_(y_min, o.y_min), fmin_(z_min, o.z_min),
                fmax_(x_max, o.x_max), fmax_(y_max, o.y_max), fmax_(z_max, o.z_max));
  }
  bool contains(const AABB& o) const {
    return o.x_min >= x_min && o.y_min >= y_min && o.z_min >= z_min &&
           o.x_max <= x_max && o.y_max <= y_max && o.z_max <= z_max;
  }
  // aabb.rs:132-164. Returns false for a miss; *t = tmin (outside) or 0 (inside).
  bool hit(const Ray& ray, float* t) const {
    float tx1 = (x_min - ray.origin.x) * ray.inv_dir.x;
    float tx2 = (x_max - ray.origin.x) * ray.inv_dir.x;
    float ty1 = (y_min - ray.origin.y) * ray.inv_dir.y;
    float ty2 = (y_max - ray.origin.y) * ray.inv_dir.y;
    float tz1 = (z_min - ray.origin.z) * ray.inv_dir.z;
    float tz2 = (z_max - ray.origin.z) * ray.inv_dir.z;
    float txmin = fmin_(tx1, tx2), tymin = fmin_(ty1, ty2), tzmin = fmin_(tz1, tz2);
    float txmax = fmax_(tx1, tx2), tymax = fmax_(ty1, ty2), tzmax = fmax_(tz1, tz2);
    float tmin = fmax_(fmax_(txmin, tymin), tzmin);
    float tmax = fmin_(fmin_(txmax, tymax), tzmax);
    if (tmin > tmax) return false;
    if (tmin >= 0.0f) { *t = tmin; return true; }
    if (tmax >= 0.0f) { *t = 0.0f; return true; }
    return false;
  }
  // One lane of AABBx4::hit, aabb.rs:252-288: -inf for a miss.
  float hit_x4_lane(const Ray& ray) const {
    float tx1 = (x_min - ray.origin.x) * ray.inv_dir.x;
    float tx2 = (x_max - ray.origin.x) * ray.inv_dir.x;
    float ty1 = (y_min - ray.origin.y) * ray.inv_dir.y;
    float ty2 = (y_max - ray.origin.y) * ray.inv_dir.y;
    float tz1 = (z_min - ray.origin.z) * ray.inv_dir.z;
    float tz2 = (z_max - ray.origin.z) * ray.inv_dir.z;
    float txmin = fmin_(tx1, tx2), tymin = fmin_(ty1, ty2), tzmin = fmin_(tz1, tz2);
    float txmax = fmax_(tx1, tx2), tymax = fmax_(ty1, ty2), tzmax = fmax_(tz1, tz2);
    float tmin = fmax_(fmax_(txmin, tymin), tzmin);
    float tmax = fmin_(fmin_(txmax, tymax), tzmax);
    if (tmin > tmax || tmax < 0.0f) return -INF_F;
    if (tmin >= 0.0f) return tmin;
    return 0.0f;
  }
};

}  // namespace ref
