// Device-side geometry, RNG and traversal. Compiled with -fmad=false and without
// -use_fast_math: every f32 expression is written in the reference's evaluation order so
// that results are bit-identical to Rust's (no FMA contraction, IEEE div/sqrt).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "wpt_types.h"

namespace wpt {

#define WPT_DEV __device__ __forceinline__
#define WPT_STACK 64          // (node, entry distance) entries; host checks BVH depth against it
#define WPT_INF CUDART_INF_F
// -DWPT_CHECKED: device-side bounds checks that trap (compute-sanitizer is closed on the pool; profiles/r2_checks.md)
#ifdef WPT_CHECKED
#define WPT_CHECK(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define WPT_CHECK(cond) do { } while (0)
#endif
// kernel variants by scene content: triangles + planes only / the reference's primitives and materials (+ torus, box) /
// everything incl. the extension (sphere, square, textures, reflect / refract — DESIGN.md 9)
enum : int { K_SIMPLE = 0, K_REF = 1, K_EXT = 2 };

// ------------------------------------------------------------------ vectors (math/vec3.rs)
struct F3 { float x, y, z; };
WPT_DEV F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
WPT_DEV F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
WPT_DEV F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
WPT_DEV F3 operator-(F3 a) { return f3(-a.x, -a.y, -a.z); }
WPT_DEV F3 operator*(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
WPT_DEV F3 operator*(F3 a, float m) { return f3(m * a.x, m * a.y, m * a.z); }   // vec3.rs:151-165
WPT_DEV F3 operator*(float m, F3 a) { return f3(m * a.x, m * a.y, m * a.z); }
WPT_DEV F3 operator/(F3 a, float d) { return f3(a.x / d, a.y / d, a.z / d); }
WPT_DEV float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
WPT_DEV F3 cross(F3 a, F3 t) { return f3(a.y * t.z - a.z * t.y, a.z * t.x - a.x * t.z, a.x * t.y - a.y * t.x); }
WPT_DEV float len(F3 a) { return sqrtf(dot(a, a)); }
WPT_DEV F3 normalize(F3 a) { return a * (1.0f / len(a)); }   // vec3.rs:27-29
WPT_DEV F3 orthogonal(F3 s) {                                 // vec3.rs:37-54
  // the three cases differ only in which component is divided out: one division and one normalize on selected
  // operands — the same operations on the same values as the reference's three branches, a third of the code
  const bool cz = fabsf(s.z) > 0.1f, cx = !cz && fabsf(s.x) > 0.1f;
  const float a = cz ? s.x : (cx ? s.y : s.x);
  const float b = cz ? s.y : s.z;
  const float den = cz ? s.z : (cx ? s.x : s.y);
  const float q = -(a * 1.0f + b * 1.0f) / den;
  return normalize(f3(cx ? q : 1.0f, (cz || cx) ? 1.0f : q, cz ? q : 1.0f));
}
WPT_DEV F3 xyz(float4 v) { return f3(v.x, v.y, v.z); }

struct Ray { F3 o, d, inv; };
WPT_DEV Ray make_ray(F3 o, F3 d) { Ray r; r.o = o; r.d = d; r.inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); return r; }   // ray.rs:31-33

// ------------------------------------------------------------------ RNG (rng.rs) + stream contract
struct Rng {
  uint32_t s;
  WPT_DEV uint32_t u32() { uint32_t x = s; x ^= x << 13; x ^= x >> 17; x ^= x << 5; s = x; return x; }
  WPT_DEV float f32() { return (float)u32() * (1.0f / 4294967296.0f); }   // rng.rs:19-21
  WPT_DEV uint32_t range(uint32_t lo, uint32_t hi) {                       // rng.rs:25-38
    if (hi == lo + 1) return 0;
    float f = f32();
    if (f == 1.0f) return hi - 1;
    return (uint32_t)floorf(f * (float)(hi - lo)) + lo;
  }
};
WPT_DEV uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
WPT_DEV uint32_t stream_seed(uint32_t index, uint32_t sample, uint32_t stream, uint32_t base) {
  uint32_t s = mix32(index + mix32(sample + mix32(stream ^ base)));
  return s == 0 ? 0xBABABEBEu : s;
}
// cos/sin of a in [0, 2*pi] from f32 + - * only (DESIGN.md "shared trig")
WPT_DEV void shared_sincos(float a, float* s_out, float* c_out) {
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

// EXTENSION: e^-x for x >= 0 from f32 + - * and exponent bits only (same code as the oracle's)
WPT_DEV float shared_exp_neg(float x) {
  if (!(x < 87.0f)) return 0.0f;
  float kf = (float)(int)(x * 1.44269504f + 0.5f);
  float r = (kf * 0.693145751953125f - x) + kf * 1.42860682030941723212e-6f;
  float p = 1.0f + r * (1.0f + r * (0.5f + r * (0.16666667f + r * (0.041666668f + r * (0.008333334f + r * 0.0013888889f)))));
  return p * __uint_as_float((uint32_t)(127 - (int)kf) << 23);
}

// ------------------------------------------------------------------ boxes (aabb.rs)
// AABB::hit, aabb.rs:132-164. Box = (a.x,a.y,a.z)-(a.w,b.x,b.y).
WPT_DEV bool box_hit(float x0, float y0, float z0, float x1, float y1, float z1, const Ray& r, float* t) {
  float tx1 = (x0 - r.o.x) * r.inv.x, tx2 = (x1 - r.o.x) * r.inv.x;
  float ty1 = (y0 - r.o.y) * r.inv.y, ty2 = (y1 - r.o.y) * r.inv.y;
  float tz1 = (z0 - r.o.z) * r.inv.z, tz2 = (z1 - r.o.z) * r.inv.z;
  float tmin = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fminf(tz1, tz2));
  float tmax = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fmaxf(tz1, tz2));
  if (tmin > tmax) return false;
  if (tmin >= 0.0f) { *t = tmin; return true; }
  if (tmax >= 0.0f) { *t = 0.0f; return true; }
  return false;
}
// one lane of AABBx4::hit, aabb.rs:252-288: -inf for a miss
WPT_DEV float box_hit_x4(float x0, float y0, float z0, float x1, float y1, float z1, const Ray& r) {
  float tx1 = (x0 - r.o.x) * r.inv.x, tx2 = (x1 - r.o.x) * r.inv.x;
  float ty1 = (y0 - r.o.y) * r.inv.y, ty2 = (y1 - r.o.y) * r.inv.y;
  float tz1 = (z0 - r.o.z) * r.inv.z, tz2 = (z1 - r.o.z) * r.inv.z;
  float tmin = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fminf(tz1, tz2));
  float tmax = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fmaxf(tz1, tz2));
  if (tmin > tmax || tmax < 0.0f) return -WPT_INF;
  if (tmin >= 0.0f) return tmin;
  return 0.0f;
}

// ------------------------------------------------------------------ roots 0.0.4 quartic (f64)
// The solver is rarely-executed code that has to stay small (instruction cache): one out-of-line copy of every f64
// library routine, division and square root it uses — the same routines, hence the same bits, a third of the code.
static __device__ __noinline__ double nl_div(double a, double b) { return a / b; }
static __device__ __noinline__ double nl_sqrt(double a) { return sqrt(a); }
#define F64_FN static __device__ __noinline__
#define F64_BITS(a) ((unsigned long long)__double_as_longlong(a))
#define F64_FROM_BITS(b) __longlong_as_double((long long)(b))

// Shared f64 acos / cos / cbrt for the torus' quartic solver (deviation B11): glibc and CUDA round these
// routines differently in the last place, and a last-place difference in f64 occasionally flips the f32 rounding of a hit
// distance (a handful of the 300 000 museum photons). Both sides therefore evaluate the same f64 + - * / sqrt sequence
// (no FMA contraction on either side): Cody-Waite reduction + Taylor kernels, the asin series, Halley iterations.
// Accuracy: <= 2 ulp against libm (tests/test_oracle_kat.py).
F64_FN double shared_cos64(double x) {
  const double kd = floor(x * 0.6366197723675814 + 0.5);
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
  if (x > 0.5) return 2.0 * shared_asin_small64(sqrt((1.0 - x) * 0.5));
  if (x < -0.5) return 3.141592653589793 - 2.0 * shared_asin_small64(sqrt((1.0 + x) * 0.5));
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
static __device__ __forceinline__ double nl_cos(double a) { return shared_cos64(a); }
static __device__ __forceinline__ double nl_acos(double a) { return shared_acos64(a); }
static __device__ __forceinline__ double nl_cbrt(double a) { return shared_cbrt64(a); }
struct Roots4 {
  int n; double v[4];
  WPT_DEV void add(double x) {
    for (int i = 0; i < n; i++) if (v[i] == x) return;
    if (n == 4) return;
    int i = n;
    while (i > 0 && v[i - 1] > x) { v[i] = v[i - 1]; i--; }
    v[i] = x; n++;
  }
};
static __device__ __noinline__ void d_quadratic_normalized(double a1, double a0, Roots4& r) {
  double disc = a1 * a1 - 4.0 * a0;
  if (disc < 0.0) return;
  double h = a1 / 2.0;
  if (disc == 0.0) { r.add(-h); return; }
  double sq = nl_sqrt(disc);
  r.add(-h - sq / 2.0);
  r.add(-h + sq / 2.0);
}
static __device__ __noinline__ void d_quadratic(double a2, double a1, double a0, Roots4& r) {
  if (a2 == 0.0) { if (a1 != 0.0) r.add(nl_div(-a0, a1)); return; }
  double disc = a1 * a1 - 4.0 * a2 * a0;
  if (disc < 0.0) return;
  double a2x2 = 2.0 * a2;
  if (disc == 0.0) { r.add(nl_div(-a1, a2x2)); return; }
  double sq = nl_sqrt(disc);
  r.add(nl_div(-a1 - sq, a2x2));
  r.add(nl_div(-a1 + sq, a2x2));
}
static __device__ __noinline__ void d_cubic_normalized(double a2, double a1, double a0, Roots4& out) {
  double q = nl_div(3.0 * a1 - a2 * a2, 9.0);
  double r = nl_div(9.0 * a2 * a1 - 27.0 * a0 - 2.0 * a2 * a2 * a2, 54.0);
  double q3 = q * q * q;
  double d = q3 + r * r;
  double a2_div_3 = nl_div(a2, 3.0);
  if (d < 0.0) {
    double phi_3 = nl_div(nl_acos(nl_div(r, nl_sqrt(-q3))), 3.0);
    double sqrt_q_2 = 2.0 * nl_sqrt(-q);
    const double two_third_pi = 2.0943951023931954923;
    out.add(sqrt_q_2 * nl_cos(phi_3) - a2_div_3);
    out.add(sqrt_q_2 * nl_cos(phi_3 - two_third_pi) - a2_div_3);
    out.add(sqrt_q_2 * nl_cos(phi_3 + two_third_pi) - a2_div_3);
  } else {
    double sqrt_d = nl_sqrt(d);
    double s = nl_cbrt(r + sqrt_d);
    double t = nl_cbrt(r - sqrt_d);
    out.add(s + t - a2_div_3);
    if (s == t && s + t != 0.0) out.add(-(s + t) / 2.0 - a2_div_3);
  }
}
static __device__ __noinline__ void d_cubic(double a3, double a2, double a1, double a0, Roots4& r) {
  if (a3 == 0.0) { d_quadratic(a2, a1, a0, r); return; }
  if (a2 == 0.0 && a1 == 0.0 && a0 == 0.0) { r.add(0.0); return; }
  d_cubic_normalized(nl_div(a2, a3), nl_div(a1, a3), nl_div(a0, a3), r);
}
static __device__ __noinline__ void d_biquadratic(double a4, double a2, double a0, Roots4& out) {
  Roots4 q; q.n = 0;
  d_quadratic(a4, a2, a0, q);
  for (int i = 0; i < q.n; i++) {
    double x = q.v[i];
    if (x > 0.0) { double s = nl_sqrt(x); out.add(-s); out.add(s); }
    else if (x == 0.0) out.add(0.0);
  }
}
static __device__ __noinline__ void d_quartic_depressed(double a2, double a1, double a0, Roots4& out) {
  if (a1 == 0.0) { d_biquadratic(1.0, a2, a0, out); return; }
  if (a0 == 0.0) { d_cubic_normalized(0.0, a2, a1, out); out.add(0.0); return; }
  double a2_pow_2 = a2 * a2;
  double a1_div_2 = a1 / 2.0;
  double b2 = a2 * 5.0 / 2.0;
  double b1 = 2.0 * a2_pow_2 - a0;
  double b0 = (a2_pow_2 * a2 - a2 * a0 - a1_div_2 * a1_div_2) / 2.0;
  Roots4 res; res.n = 0;
  d_cubic_normalized(b2, b1, b0, res);
  double y = res.v[res.n - 1];
  double a2_plus_2y = a2 + 2.0 * y;
  if (a2_plus_2y > 0.0) {
    double s = nl_sqrt(a2_plus_2y);
    double q0a = a2 + y - nl_div(a1_div_2, s);
    double q0b = a2 + y + nl_div(a1_div_2, s);
    Roots4 ra; ra.n = 0; Roots4 rb; rb.n = 0;
    d_quadratic_normalized(s, q0a, ra);
    d_quadratic_normalized(-s, q0b, rb);
    for (int i = 0; i < ra.n; i++) out.add(ra.v[i]);
    for (int i = 0; i < rb.n; i++) out.add(rb.v[i]);
  }
}
static __device__ __noinline__ void d_quartic(double a4, double a3, double a2, double a1, double a0, Roots4& out) {
  out.n = 0;
  if (a4 == 0.0) { d_cubic(a3, a2, a1, a0, out); return; }
  if (a0 == 0.0) { d_cubic(a4, a3, a2, a1, out); out.add(0.0); return; }
  if (a1 == 0.0 && a3 == 0.0) { d_biquadratic(a4, a2, a0, out); return; }
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
  if (discriminant == 0.0) {
    bool triple = delta0 == 0.0;
    bool quadruple = triple && dd == 0.0;
    bool no_roots = dd == 0.0 && pp > 0.0 && rr == 0.0;
    if (quadruple) { out.add(nl_div(-a3, 4.0 * a4)); return; }
    if (triple) {
      double x0 = nl_div(-72.0 * a4 * a4 * a0 + 10.0 * a4 * a2 * a2 - 3.0 * a3 * a3 * a2,
                         9.0 * (8.0 * a4 * a4 * a1 - 4.0 * a4 * a3 * a2 + a3 * a3 * a3));
      out.add(x0);
      out.add(-(nl_div(a3, a4) + 3.0 * x0));
      return;
    }
    if (no_roots) return;
  } else if (discriminant > 0.0 && (pp > 0.0 || dd > 0.0)) return;
  double a4_pow_2 = a4 * a4, a4_pow_3 = a4_pow_2 * a4, a4_pow_4 = a4_pow_2 * a4_pow_2;
  double p = nl_div(pp, 8.0 * a4_pow_2);
  double q = nl_div(rr, 8.0 * a4_pow_3);
  double r = nl_div(dd + 16.0 * a4_pow_2 * (12.0 * a0 * a4 - 3.0 * a1 * a3 + a2 * a2), 256.0 * a4_pow_4);
  Roots4 dep; dep.n = 0;
  d_quartic_depressed(p, q, r, dep);
  const double shift = nl_div(a3, 4.0 * a4);
  for (int i = 0; i < dep.n; i++) out.add(dep.v[i] - shift);
}

// ------------------------------------------------------------------ primitives
// Conservative f32 cull in front of the f64 quartic (no reference counterpart: it only skips solver calls that cannot
// return a root). The flat torus (axis y, torus.rs:11-16) lies inside the slab |y - c.y| <= r and between the cylinders of
// radius R - r and R + r around its axis. Over the part of the ray inside the slab (widened by the margin m) the squared
// distance from the axis is a convex quadratic in t: if its maximum there is below (R - r - m)^2 the ray passes through the
// hole, if its minimum is above (R + r + m)^2 it passes outside. m = 0.02 is four orders of magnitude above the f32
// rounding of these expressions, and a ray that stays 0.02 away from the surface has no real root (F >= (2 r m)^2).
// Only for origins within 64 units of the torus: from tens of thousands of units away (grazing bounces off the infinite
// floor) the reference's f64 quartic itself returns numerically spurious roots, which have to be reproduced, not fixed.
WPT_DEV bool torus_may_hit(float4 q0, float4 q1, const Ray& ray) {
  const float R = q1.x, r = q1.y, m = 0.02f;
  const float ox = ray.o.x - q0.x, oy = ray.o.y - q0.y, oz = ray.o.z - q0.z;
  if (!(fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz)) <= 64.0f)) return true;
  const float hy = r + m;
  float t_lo = 0.0f, t_hi = WPT_INF;
  if (fabsf(ray.d.y) > 1e-6f) {
    const float ta = (-hy - oy) * ray.inv.y, tb = (hy - oy) * ray.inv.y;
    t_lo = fmaxf(fminf(ta, tb), 0.0f); t_hi = fmaxf(ta, tb);
    if (!(t_hi >= t_lo)) return false;            // the slab lies behind the origin
  } else if (fabsf(oy) > hy) return false;        // parallel to the slab and outside of it
  const float a = ray.d.x * ray.d.x + ray.d.z * ray.d.z, b = ox * ray.d.x + oz * ray.d.z, c = ox * ox + oz * oz;
  const float outer = (R + r + m) * (R + r + m);
  float t_min = a > 0.0f ? fminf(fmaxf(-b / a, t_lo), t_hi) : t_lo;   // the vertex clamped into the interval (t_hi may be inf)
  if (!(t_min <= 3.0e38f)) t_min = t_lo;
  const float rho_min = (a * t_min + 2.0f * b) * t_min + c;
  if (rho_min > outer) return false;
  if (t_hi <= 3.0e38f) {
    const float inner = fmaxf(R - r - m, 0.0f) * fmaxf(R - r - m, 0.0f);
    const float rho_a = (a * t_lo + 2.0f * b) * t_lo + c, rho_b = (a * t_hi + 2.0f * b) * t_hi + c;
    if (fmaxf(rho_a, rho_b) < inner) return false;
  }
  return true;
}
// Torus::trace, torus.rs:61-126. Returns hit distance (f32) and outward-or-flipped normal.
static __device__ __noinline__ bool torus_trace(float4 q0, float4 q1, const Ray& ray, float* t_out, F3* n_out, bool* entering_out = nullptr) {
  double a = (double)q1.x, b = (double)q1.y;
  F3 d = ray.o - xyz(q0);
  F3 e = ray.d;
  double dx = d.x, dy = d.y, dz = d.z, ex = e.x, ey = e.y, ez = e.z;
  double g = 4.0 * a * a * (ex * ex + ez * ez);
  double h = 8.0 * a * a * (dx * ex + dz * ez);
  double i = 4.0 * a * a * (dx * dx + dz * dz);
  double j = ex * ex + ey * ey + ez * ez;
  double k = 2.0 * (dx * ex + dy * ey + dz * ez);
  double l = dx * dx + dy * dy + dz * dz + a * a - b * b;
  Roots4 rs;
  d_quartic(j * j, 2.0 * j * k, 2.0 * j * l + k * k - g, 2.0 * k * l - h, l * l - i, rs);
  int np = 0; double closest = 0.0;
  for (int q = 0; q < rs.n; q++)
    if (rs.v[q] >= 0.0001) { closest = np == 0 ? rs.v[q] : fmin(closest, rs.v[q]); np++; }   // torus.rs:130-139,107-110
  if (np == 0) return false;
  *t_out = (float)closest;
  if (n_out) {
    double px = (double)d.x + (double)e.x * closest;
    double py = (double)d.y + (double)e.y * closest;
    double pz = (double)d.z + (double)e.z * closest;
    double alpha = 1.0 - nl_div(a, nl_sqrt(px * px + pz * pz));
    F3 n = normalize(f3((float)(alpha * px), (float)py, (float)(alpha * pz)));
    *n_out = (np % 2 == 1) ? -n : n;   // odd number of positive roots: inside (torus.rs:120-124)
    if (entering_out) *entering_out = (np % 2 == 0);
  }
  return true;
}

// Sphere (sphere.rs:55-82, 106-128): algebraic solution with a = 1
WPT_DEV bool sphere_t(float4 q0, float4 q1, const Ray& ray, float* t_out, bool* entering) {
  F3 oc = ray.o - xyz(q0);
  float a = 1.0f;
  float b = 2.0f * dot(ray.d, oc);
  float c = dot(oc, oc) - q1.x * q1.x;
  float d = b * b - 4.0f * a * c;
  if (d < 0.0f) return false;
  float ds = sqrtf(d);
  float t0 = (-b + ds) / (2.0f * a);
  float t1 = (-b - ds) / (2.0f * a);
  float t = fminf(t0, t1);
  bool ent = true;
  if (t <= 0.0f) {
    t = fmaxf(t0, t1);
    if (t <= 0.0f) return false;
    ent = false;
  }
  *t_out = t; *entering = ent;
  return true;
}
// Square (square.rs:56-99): finite upward-facing plane; uv as the reference computes it for a textured material
WPT_DEV bool square_t(float4 q0, float4 q1, const Ray& ray, float* t_out, float* u, float* v) {
  float n_dot_dir = ray.d.y;
  if (n_dot_dir == 0.0f) return false;
  float t = (q0.y - ray.o.y) / n_dot_dir;
  if (t <= 0.0f) return false;
  F3 hit = ray.o + t * ray.d;
  float dx = fabsf(hit.x - q0.x), dz = fabsf(hit.z - q0.z);
  if (2.0f * dx >= q1.x || 2.0f * dz >= q1.x) return false;
  *t_out = t;
  *u = (hit.x - q0.x) / q1.x + 0.5f; *v = (hit.z - q0.z) / q1.x + 0.5f;
  return true;
}

// Tracable::trace_simple for shape record `s`. `limit`/`strict` implement the acceptance test
// of trace_shapes_md (scene.rs:450-472): the first candidate needs t <= max_dis, later ones
// 0 < t < best. Rejecting on t before the edge tests does not change any result.
// (q0, q1 = the first two words of the record at p, already loaded; TORUS = false leaves the torus branch out: the caller
// handles tori itself — k_mega's deferred solver phase)
template <int KIND, bool TORUS>
WPT_DEV bool shape_trace_simple_q(const float4* __restrict__ p, float4 q0, float4 q1, const Ray& ray, float limit, bool strict, float* t_out) {
  uint32_t type = __float_as_uint(q0.w) & 0xFFu;
  if (type == SH_TRIANGLE) {   // triangle.rs:159-191
    float4 q2 = __ldg(p + 2), q3 = __ldg(p + 3);
    F3 v0 = xyz(q0), v1 = xyz(q1), v2 = xyz(q2);
    F3 n = f3(q1.w, q2.w, q3.x);
    float n_dot_d = dot(n, ray.d);
    if (n_dot_d == 0.0f) return false;
    float orig_dis = dot(n, v0);
    float t = (orig_dis - dot(n, ray.o)) / n_dot_d;
    if (t <= 0.0f) return false;
    if (strict ? !(t < limit) : !(t <= limit)) return false;
    F3 nn = f3(q3.y, q3.z, q3.w);
    F3 pt = ray.o + t * ray.d;
    const float slack = 0.1f * WPT_EPSILON;   // triangle.rs:41-45
    if (!(dot(nn, cross(v1 - v0, pt - v0)) + slack >= 0.0f)) return false;
    if (!(dot(nn, cross(v2 - v1, pt - v1)) + slack >= 0.0f)) return false;
    if (!(dot(nn, cross(v0 - v2, pt - v2)) + slack >= 0.0f)) return false;
    *t_out = t;
    return true;
  }
  float t;
  if (KIND == K_SIMPLE || type == SH_PLANE) {      // plane.rs:80-99 (K_SIMPLE scenes hold triangles and planes only)
    F3 nr = xyz(q1);
    float n_dot_dir = dot(nr, ray.d);
    if (n_dot_dir == 0.0f) return false;
    t = (q1.w - dot(nr, ray.o)) / n_dot_dir;
    if (t <= 0.0f) return false;
  } else if (type == SH_AARECT) {   // aa_rect.rs:142-174 (own 1/dir: the same value as ray.inv)
    float tx1 = (q0.x - ray.o.x) * ray.inv.x, tx2 = (q1.x - ray.o.x) * ray.inv.x;
    float ty1 = (q0.y - ray.o.y) * ray.inv.y, ty2 = (q1.y - ray.o.y) * ray.inv.y;
    float tz1 = (q0.z - ray.o.z) * ray.inv.z, tz2 = (q1.z - ray.o.z) * ray.inv.z;
    float tmin = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fminf(tz1, tz2));
    float tmax = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fmaxf(tz1, tz2));
    if (tmin >= tmax) return false;
    if (tmin > 0.0f) t = tmin;
    else if (tmax > 0.0f) t = tmax;
    else return false;
  } else if (KIND == K_EXT && type == SH_SPHERE) {
    bool ent;
    if (!sphere_t(q0, q1, ray, &t, &ent)) return false;
  } else if (KIND == K_EXT && type == SH_SQUARE) {   // ray.rs:110-116 default = trace().distance
    float u, v;
    if (!square_t(q0, q1, ray, &t, &u, &v)) return false;
  } else if (TORUS) {          // torus: ray.rs:110-116 default = trace().distance
    if (!torus_may_hit(q0, q1, ray) || !torus_trace(q0, q1, ray, &t, nullptr)) return false;
  } else return false;
  if (strict ? !(0.0f < t && t < limit) : !(t <= limit)) return false;
  *t_out = t;
  return true;
}
template <int KIND>
WPT_DEV bool shape_trace_simple(const DShape* __restrict__ shapes, uint32_t idx, const Ray& ray, float limit, bool strict, float* t_out) {
  const float4* p = reinterpret_cast<const float4*>(shapes + idx);
  float4 q0 = __ldg(p), q1 = __ldg(p + 1);
  return shape_trace_simple_q<KIND, true>(p, q0, q1, ray, limit, strict, t_out);
}

// Tracable::trace for the winning shape (scene.rs:140): distance, Hit::new-normalised normal,
// material index. Returns false if the full intersection reports no hit.
// PRE (k_mega with the torus phase): a torus winner's distance / normal / is_entering were computed by trav_torus (t_pre, pre).
struct TorusPre { F3 n; bool entering; };
WPT_DEV void torus_pre_get(const TorusPre* pre, F3* n, bool* entering) { *n = pre->n; *entering = pre->entering; }
template <int KIND, bool PRE = false>
WPT_DEV bool shape_trace_full(const DShape* __restrict__ shapes, uint32_t idx, const Ray& ray, float* t_out, F3* n_out, uint32_t* mat_out, bool* entering_out = nullptr, float2* uv_out = nullptr,
                              float t_pre = 0.0f, const TorusPre* pre = nullptr) {
  const float4* p = reinterpret_cast<const float4*>(shapes + idx);
  float4 q0 = __ldg(p), q1 = __ldg(p + 1);
  uint32_t meta = __float_as_uint(q0.w);
  uint32_t type = meta & 0xFFu;
  *mat_out = meta >> 8;
  F3 n; float t;
  bool entering = true;   // Hit::is_entering (only the extension's refracting material reads it)
  if (PRE && type == SH_TRIANGLE) {   // the traversal's distance (see shape_hit_normal_tri_plane): only the side test is left
    float4 q2 = __ldg(p + 2), q3 = __ldg(p + 3);
    const float n_dot_d = dot(f3(q1.w, q2.w, q3.x), ray.d);
    const F3 nn = f3(q3.y, q3.z, q3.w);
    t = t_pre; n = (n_dot_d > 0.0f) ? -nn : nn; entering = !(n_dot_d > 0.0f);
  } else if (PRE && type == SH_PLANE) {
    const F3 nr = xyz(q1);
    t = t_pre; n = (dot(nr, ray.d) > 0.0f) ? -nr : nr;
  } else if (type == SH_TRIANGLE) {   // triangle.rs:116-157
    float4 q2 = __ldg(p + 2), q3 = __ldg(p + 3);
    F3 v0 = xyz(q0), v1 = xyz(q1), v2 = xyz(q2);
    F3 nu = f3(q1.w, q2.w, q3.x);
    float n_dot_d = dot(nu, ray.d);
    if (n_dot_d == 0.0f) return false;
    t = (dot(nu, v0) - dot(nu, ray.o)) / n_dot_d;
    if (t <= 0.0f) return false;
    F3 nn = f3(q3.y, q3.z, q3.w);
    F3 pt = ray.o + t * ray.d;
    const float slack = 0.1f * WPT_EPSILON;
    if (!(dot(nn, cross(v1 - v0, pt - v0)) + slack >= 0.0f)) return false;
    if (!(dot(nn, cross(v2 - v1, pt - v1)) + slack >= 0.0f)) return false;
    if (!(dot(nn, cross(v0 - v2, pt - v2)) + slack >= 0.0f)) return false;
    n = (n_dot_d > 0.0f) ? -nn : nn;
    entering = !(n_dot_d > 0.0f);
  } else if (KIND == K_SIMPLE || type == SH_PLANE) {   // plane.rs:45-77
    F3 nr = xyz(q1);
    float n_dot_dir = dot(nr, ray.d);
    if (n_dot_dir == 0.0f) return false;
    t = (q1.w - dot(nr, ray.o)) / n_dot_dir;
    if (t <= 0.0f) return false;
    n = (n_dot_dir > 0.0f) ? -nr : nr;
  } else if (type == SH_AARECT) {  // aa_rect.rs:71-139
    float tx1 = (q0.x - ray.o.x) * ray.inv.x, tx2 = (q1.x - ray.o.x) * ray.inv.x;
    float ty1 = (q0.y - ray.o.y) * ray.inv.y, ty2 = (q1.y - ray.o.y) * ray.inv.y;
    float tz1 = (q0.z - ray.o.z) * ray.inv.z, tz2 = (q1.z - ray.o.z) * ray.inv.z;
    float tmin = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fminf(tz1, tz2));
    float tmax = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fmaxf(tz1, tz2));
    if (tmin >= tmax) return false;
    if (tmin > 0.0f) {
      t = tmin;
      if (tmin == tx1) n = f3(-1, 0, 0); else if (tmin == tx2) n = f3(1, 0, 0);
      else if (tmin == ty1) n = f3(0, -1, 0); else if (tmin == ty2) n = f3(0, 1, 0);
      else if (tmin == tz1) n = f3(0, 0, -1); else n = f3(0, 0, 1);
    } else if (tmax > 0.0f) {
      t = tmax; entering = false;
      if (tmax == tx1) n = f3(1, 0, 0); else if (tmax == tx2) n = f3(-1, 0, 0);
      else if (tmax == ty1) n = f3(0, 1, 0); else if (tmax == ty2) n = f3(0, -1, 0);
      else if (tmax == tz1) n = f3(0, 0, 1); else n = f3(0, 0, -1);
    } else return false;
  } else if (KIND == K_EXT && type == SH_SPHERE) {   // sphere.rs:49-101
    bool ent;
    if (!sphere_t(q0, q1, ray, &t, &ent)) return false;
    n = ((ray.o + t * ray.d) - xyz(q0)) / q1.x;
    if (!ent) n = -n;
    entering = ent;
  } else if (KIND == K_EXT && type == SH_SQUARE) {   // square.rs:56-99
    float u, v;
    if (!square_t(q0, q1, ray, &t, &u, &v)) return false;
    n = ray.d.y > 0.0f ? f3(0, -1, 0) : f3(0, 1, 0);
    if (uv_out) *uv_out = make_float2(u, v);
  } else if (PRE) {
    torus_pre_get(pre, &n, &entering); t = t_pre;
  } else {
    if (!torus_may_hit(q0, q1, ray) || !torus_trace(q0, q1, ray, &t, &n, &entering)) return false;
  }
  *t_out = t;
  *n_out = normalize(n);   // Hit::new, ray.rs:59-62
  if (entering_out) *entering_out = entering;
  return true;
}

// Normal + material of a hit whose distance `t` is already known from the traversal. Scene::trace
// re-intersects the winner with Tracable::trace (scene.rs:140); for triangles and planes that
// recomputes exactly the same `t` and, since the shape was just hit, cannot miss — only the side
// test n.d > 0 and the Hit::new normalisation are left (triangle.rs:143-147, plane.rs:61-75).
WPT_DEV void shape_hit_normal_tri_plane(const DShape* __restrict__ shapes, uint32_t idx, const Ray& ray, F3* n_out, uint32_t* mat_out) {
  const float4* p = reinterpret_cast<const float4*>(shapes + idx);
  float4 q0 = __ldg(p), q1 = __ldg(p + 1);
  uint32_t meta = __float_as_uint(q0.w);
  *mat_out = meta >> 8;
  F3 n;
  if ((meta & 0xFFu) == SH_TRIANGLE) {
    float4 q2 = __ldg(p + 2), q3 = __ldg(p + 3);
    F3 nu = f3(q1.w, q2.w, q3.x);
    F3 nn = f3(q3.y, q3.z, q3.w);
    n = (dot(nu, ray.d) > 0.0f) ? -nn : nn;
  } else {
    F3 nr = xyz(q1);
    n = (dot(nr, ray.d) > 0.0f) ? -nr : nr;
  }
  *n_out = normalize(n);   // Hit::new, ray.rs:59-62
}

// ------------------------------------------------------------------ Scene::trace_g (scene.rs:162-184)
struct GHit { float t; int id; uint32_t visits; uint32_t prims; };

// trace_shapes_md over a leaf (scene.rs:450-472); updates (best_t, best_id) if the leaf
// reports a hit — a later leaf wins exact ties because the test is t <= max_dis.
// (Tried in round 2 and removed: intersecting a leaf's tori first, with all lanes of the leaf step that hold one, and taking
//  the distances from a cache in the ordered scan. The solver then starts with 14.5 lanes instead of 3, but diverges inside
//  on the discriminant cases, and the extra loops cost more than they save: 55.5 ms against 54.0 ms on the museum frame,
//  gpurun_out/r2_ab_museum3.log.)
template <int KIND>
WPT_DEV void leaf_scan(const DScene& sc, uint32_t first, uint32_t count, const Ray& ray, float& bound, int& best_id, uint32_t& prims) {
  prims += count;
  WPT_CHECK(first + count <= sc.num_shapes);
  bool have = false; float bt = 0.0f; uint32_t bi = 0;
  for (uint32_t i = 0; i < count; i++) {
    float t;
    if (shape_trace_simple<KIND>(sc.shapes, first + i, ray, have ? bt : bound, have, &t)) { have = true; bt = t; bi = first + i; }
  }
  if (have) { bound = bt; best_id = (int)bi; }
}

// ------------------------------------------------------------------ resumable traversal
// Scene::trace_g split into begin / step / result so that a persistent kernel can interleave
// the traversal of different lanes (k_mega) while the wavefront kernel (k_trace) simply loops.
//
// BVH2: traverse_bvh_guarded + traverse_bvh (scene.rs:191-288), iterative with an explicit
// stack of (node, entry distance): a popped node is skipped iff the current bound is < its
// entry distance — exactly the `lshape_dis < right_dis` early-outs of the recursion.
// BVH4: traverse_bvh4 (scene.rs:292-342): children are pushed in reverse sorted order; a popped
// child is dropped iff its box distance is > the current bound, which is what the recursion's
// early `return` does for it and for all later (farther) siblings.
struct Trav {
  uint32_t lf, cnt;      // BVH2: record of the node to enter next; BVH4: lf = node id / leaf code
  int sp;                // stack size
  float bound;           // min(plane hit, best BVH hit so far) = the recursion's max_dis
  int best_id;           // best BVH hit (-1: none)
  float inf_t; int inf_id;   // hit among the infinite shapes (scene.rs:168,176)
  uint32_t visits, prims;
  uint32_t lcur;         // k_mega with the deferred torus phase: leaf-scan state, cursor | have << 8 | resumed << 9 (0 = not inside a leaf)
};

// scene.rs:346-388 — the exact compare-and-swap network (not stable for n == 4).
// -DWPT_SORT_PREDICATED: the same network with predicated swaps — every swap of the reference happens iff the branch it sits in is
// taken and its comparison (on the current values) holds, so the same permutation comes out, ties and NaNs included (checked
// exhaustively over 7^4 x 5 inputs), with no divergent branch. Measured and not used: BVH4 + PNEE 30.7 -> 30.6 ms, BVH4 NEE
// 22.3 -> 22.5 ms, NoNEE 15.6 -> 15.8 ms (gpurun_out/r2n_ab.log) — the branches ran with 5 – 7 lanes, the selects run for everyone.
#ifdef WPT_SORT_PREDICATED
WPT_DEV void sort_small(int* id, float* d, uint32_t n) {
#define WPT_CSWAP(i, j, c) { const bool c_ = (c); const int ti = id[i], tj = id[j]; const float di = d[i], dj = d[j]; id[i] = c_ ? tj : ti; id[j] = c_ ? ti : tj; d[i] = c_ ? dj : di; d[j] = c_ ? di : dj; }
  const bool n2 = n >= 2, n3 = n == 3, n4 = n == 4;
  WPT_CSWAP(0, 1, n2 && d[1] < d[0])          // first swap of all three networks
  // n == 3: (1,2), (0,1)
  WPT_CSWAP(1, 2, n3 && d[2] < d[1])
  WPT_CSWAP(0, 1, n3 && d[1] < d[0])
  // n == 4
  WPT_CSWAP(2, 3, n4 && d[3] < d[2])
  const bool A = n4 && d[0] < d[2], B = n4 && !(d[0] < d[2]);
  const bool a1 = A && d[2] < d[1];
  WPT_CSWAP(1, 2, a1)
  WPT_CSWAP(2, 3, a1 && d[3] < d[2])
  WPT_CSWAP(0, 2, B)
  WPT_CSWAP(1, 2, B)
  const bool b1 = B && d[3] < d[1];
  const bool b2 = B && !b1 && d[3] < d[2];
  WPT_CSWAP(1, 3, b1)
  WPT_CSWAP(2, 3, b1 || b2)
#undef WPT_CSWAP
}
#else
WPT_DEV void sort_small(int* id, float* d, uint32_t n) {
#define WPT_SWAP(i, j) { int ti = id[i]; id[i] = id[j]; id[j] = ti; float td = d[i]; d[i] = d[j]; d[j] = td; }
  if (n == 2) {
    if (d[1] < d[0]) WPT_SWAP(0, 1)
  } else if (n == 3) {
    if (d[1] < d[0]) WPT_SWAP(0, 1)
    if (d[2] < d[1]) WPT_SWAP(1, 2)
    if (d[1] < d[0]) WPT_SWAP(0, 1)
  } else if (n == 4) {
    if (d[1] < d[0]) WPT_SWAP(0, 1)
    if (d[3] < d[2]) WPT_SWAP(2, 3)
    if (d[0] < d[2]) {
      if (d[2] < d[1]) {
        WPT_SWAP(1, 2)
        if (d[3] < d[2]) WPT_SWAP(2, 3)
      }
    } else {
      WPT_SWAP(0, 2)
      WPT_SWAP(1, 2)
      if (d[3] < d[1]) { WPT_SWAP(1, 3) WPT_SWAP(2, 3) }
      else if (d[3] < d[2]) WPT_SWAP(2, 3)
    }
  }
#undef WPT_SWAP
}
#endif

// trace_shapes over the infinite shapes + the root guard. Returns true if the BVH has to be
// traversed (then call trav_step until it returns false).
template <int BVH, int KIND>
WPT_DEV bool trav_begin(const DScene& sc, const Ray& ray, Trav& tv) {
  bool have = false; float it = 0.0f; int iid = -1;
  for (uint32_t i = 0; i < sc.num_inf; i++) {   // scene.rs:426-445: first hit accepted as is
    float t;
    if (shape_trace_simple<KIND>(sc.shapes, i, ray, have ? it : WPT_INF, have, &t)) { have = true; it = t; iid = (int)i; }
  }
  tv.inf_t = it; tv.inf_id = iid;
  tv.bound = have ? it : WPT_INF;
  tv.best_id = -1;
  tv.visits = 0; tv.prims = 0; tv.sp = 0; tv.lcur = 0u;
  if (BVH == 4) { tv.lf = 0u; tv.cnt = 0u; return true; }   // no root box test (scene.rs:292-342)
  const float4* __restrict__ nodes = reinterpret_cast<const float4*>(sc.nodes2);
  float4 ra = __ldg(nodes), rb = __ldg(nodes + 1);
  tv.visits = 1;   // the root guard (scene.rs:207,210)
  float h;
  if (!(box_hit(ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, ray, &h) && h < tv.bound)) return false;
  tv.lf = __float_as_uint(rb.z); tv.cnt = __float_as_uint(rb.w);
  return true;
}

// The traversal is cut into three kinds of step so that a warp can run each kind with all the
// lanes that need it (one code site each): enter an inner node, scan a leaf, pop.
// AABB::hit (aabb.rs:132-164) without early returns: same comparisons, same NaN behaviour.
WPT_DEV bool box_hit_sel(float x0, float y0, float z0, float x1, float y1, float z1, const Ray& r, float* t) {
  float tx1 = (x0 - r.o.x) * r.inv.x, tx2 = (x1 - r.o.x) * r.inv.x;
  float ty1 = (y0 - r.o.y) * r.inv.y, ty2 = (y1 - r.o.y) * r.inv.y;
  float tz1 = (z0 - r.o.z) * r.inv.z, tz2 = (z1 - r.o.z) * r.inv.z;
  float tmin = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fminf(tz1, tz2));
  float tmax = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fmaxf(tz1, tz2));
  bool front = tmin >= 0.0f;
  *t = front ? tmin : 0.0f;
  return !(tmin > tmax) && (front || tmax >= 0.0f);
}
// pop: next node to enter, skipping entries the current bound has made unreachable. Returns
// false when the stack is empty (the traversal is finished).
template <int BVH>
WPT_DEV bool trav_pop(const DScene& sc, Trav& tv, const uint32_t* stack_n, const float* stack_d) {
  if (BVH == 4) {
    while (tv.sp > 0) {
      tv.sp--;
      if (stack_d[tv.sp] > tv.bound) continue;
      tv.lf = stack_n[tv.sp];
      return true;
    }
    return false;
  }
  const float4* __restrict__ nodes = reinterpret_cast<const float4*>(sc.nodes2);
  while (tv.sp > 0) {
    tv.sp--;
    if (tv.bound < stack_d[tv.sp]) continue;   // near hit closer than the far box: skip it
    float4 nb = __ldg(nodes + (size_t)stack_n[tv.sp] * 2 + 1);
    tv.lf = __float_as_uint(nb.z); tv.cnt = __float_as_uint(nb.w);
    return true;
  }
  return false;
}
// Enter the inner node tv.lf. Returns true if the lane has to pop next (no child survived).
template <int BVH>
WPT_DEV bool trav_inner(const DScene& sc, const Ray& ray, Trav& tv, uint32_t* stack_n, float* stack_d) {
  if (BVH == 4) {
    tv.visits += 1;
    const float4* p = reinterpret_cast<const float4*>(sc.nodes4 + (int)tv.lf);
    float4 x0 = __ldg(p), y0 = __ldg(p + 1), z0 = __ldg(p + 2), x1 = __ldg(p + 3), y1 = __ldg(p + 4), z1 = __ldg(p + 5);
    int4 ch = __ldg(reinterpret_cast<const int4*>(p + 6));
    uint32_t nc = __ldg(reinterpret_cast<const uint32_t*>(p + 7));
    int id[4] = {0, 0, 0, 0}; float d[4] = {WPT_INF, WPT_INF, WPT_INF, WPT_INF};
    if (nc > 0) { id[0] = ch.x; d[0] = box_hit_x4(x0.x, y0.x, z0.x, x1.x, y1.x, z1.x, ray); }
    if (nc > 1) { id[1] = ch.y; d[1] = box_hit_x4(x0.y, y0.y, z0.y, x1.y, y1.y, z1.y, ray); }
    if (nc > 2) { id[2] = ch.z; d[2] = box_hit_x4(x0.z, y0.z, z0.z, x1.z, y1.z, z1.z, ray); }
    if (nc > 3) { id[3] = ch.w; d[3] = box_hit_x4(x0.w, y0.w, z0.w, x1.w, y1.w, z1.w, ray); }
    sort_small(id, d, nc);
    // surviving children from the farthest to the nearest: all but the nearest go on the stack, the nearest is entered
    // directly (its box distance was just tested against the bound, so the pop test could not drop it)
    int pend_id = 0; float pend_d = 0.0f; bool have = false;
#pragma unroll
    for (int i = 3; i >= 0; i--)
      if ((uint32_t)i < nc && d[i] >= 0.0f && !(d[i] > tv.bound)) {
        if (have) { WPT_CHECK(tv.sp < WPT_STACK); stack_n[tv.sp] = (uint32_t)pend_id; stack_d[tv.sp] = pend_d; tv.sp++; }
        pend_id = id[i]; pend_d = d[i]; have = true;
      }
    if (have) { tv.lf = (uint32_t)pend_id; return false; }
    return true;
  }
  // BVH2, scene.rs:241-287 written with selects instead of four branches:
  //   left misses          -> traverse_bvh_guarded(right): the guard is counted (scene.rs:283-286)
  //   only left hits       -> left
  //   both hit             -> near first (tie -> right first, scene.rs:244,261), far pushed with its distance
  const float4* __restrict__ nodes = reinterpret_cast<const float4*>(sc.nodes2);
  const float4* c = nodes + (size_t)tv.lf * 2;
  float4 la = __ldg(c), lb = __ldg(c + 1), qa = __ldg(c + 2), qb = __ldg(c + 3);
  float dl, dr;
  bool hl = box_hit_sel(la.x, la.y, la.z, la.w, lb.x, lb.y, ray, &dl) && dl < tv.bound;
  bool hr = box_hit_sel(qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, ray, &dr) && dr < tv.bound;
  tv.visits += hl ? 1u : 2u;
  const bool lfirst = dl < dr;
  const bool go_left = hl && (!hr || lfirst);
  if (hl && hr) { WPT_CHECK(tv.sp < WPT_STACK); stack_n[tv.sp] = lfirst ? tv.lf + 1 : tv.lf; stack_d[tv.sp] = lfirst ? dr : dl; tv.sp++; }
  tv.lf = __float_as_uint(go_left ? lb.z : qb.z); tv.cnt = __float_as_uint(go_left ? lb.w : qb.w);
  return !(hl || hr);
}
// Scan the leaf the lane waits at (the lane pops afterwards).
template <int BVH, int KIND>
WPT_DEV void trav_leaf(const DScene& sc, const Ray& ray, Trav& tv) {
  tv.visits += 1;
  if (BVH == 4) { uint32_t code = tv.lf; leaf_scan<KIND>(sc, sc.num_inf + (code & 0x7FFFFFFu), (code >> 27) & 0xFu, ray, tv.bound, tv.best_id, tv.prims); }
  else leaf_scan<KIND>(sc, sc.num_inf + tv.lf, tv.cnt, ray, tv.bound, tv.best_id, tv.prims);
}
// Resumable leaf scan for scenes with tori (k_mega's deferred solver phase): the same ordered scan as leaf_scan — the first
// accepted candidate of the leaf needs t <= bound, later ones 0 < t < best (scene.rs:450-472); writing an accepted candidate
// to (bound, best_id) at once is the same as leaf_scan's (bt, bi) + final copy — but a torus that passes the conservative cull
// stops the scan: the lane parks with its cursor in tv.lcur and returns true; trav_torus resolves it and the scan goes on.
enum : uint32_t { LC_HAVE = 1u << 8, LC_RESUMED = 1u << 9,   // state of the leaf scan in progress (cleared at the end of the leaf)
                  LC_BEST_TORUS = 1u << 10,                  // the best BVH hit so far is a torus (lives until the next trav_begin)
                  LC_SHADE = 1u << 11, LC_NREADY = 1u << 12, LC_ENTERING = 1u << 13 };   // normal of the winning torus: wanted / stored in (lf, cnt, sp) / Hit::is_entering
template <int BVH>
WPT_DEV void leaf_range(const DScene& sc, const Trav& tv, uint32_t* first, uint32_t* count) {
  if (BVH == 4) { *first = sc.num_inf + (tv.lf & 0x7FFFFFFu); *count = (tv.lf >> 27) & 0xFu; }
  else { *first = sc.num_inf + tv.lf; *count = tv.cnt; }
}
template <int BVH, int KIND>
WPT_DEV bool trav_leaf_deferred(const DScene& sc, const Ray& ray, Trav& tv) {
  uint32_t first, count; leaf_range<BVH>(sc, tv, &first, &count);
  const uint32_t st = tv.lcur;
  uint32_t i = st & 0xFFu; bool have = (st & LC_HAVE) != 0;
  uint32_t best_torus = st & LC_BEST_TORUS;
  if (!(st & LC_RESUMED)) { tv.visits += 1; tv.prims += count; WPT_CHECK(first + count <= sc.num_shapes); }
  for (; i < count; i++) {
    const float4* p = reinterpret_cast<const float4*>(sc.shapes + first + i);
    const float4 q0 = __ldg(p), q1 = __ldg(p + 1);
    if ((__float_as_uint(q0.w) & 0xFFu) == SH_TORUS) {
      if (!torus_may_hit(q0, q1, ray)) continue;
      tv.lcur = best_torus | i | (have ? LC_HAVE : 0u) | LC_RESUMED;
      return true;
    }
    float t;
    if (shape_trace_simple_q<KIND, false>(p, q0, q1, ray, tv.bound, have, &t)) { have = true; tv.bound = t; tv.best_id = (int)(first + i); best_torus = 0u; }
  }
  tv.lcur = best_torus;
  return false;
}
// The parked lanes of k_mega's torus phase, one solver call site for both kinds:
//  * leaf mode — the torus trav_leaf_deferred stopped at: Torus::trace (f64 quartic) + the scan's acceptance test, the cursor
//    moves on. Returns true if the leaf is finished (the lane pops next);
//  * shade mode (LC_SHADE) — the traversal is over and its winner is a torus: Scene::trace intersects the winner again for the
//    Hit (scene.rs:140); same ray, same shape, so the same distance — only the normal and is_entering are new. They are kept
//    in (lf, cnt, sp), which are dead until the next trav_begin, and shade_hit takes them from there (TorusPre). Returns false.
template <int BVH>
WPT_DEV bool trav_torus(const DScene& sc, const Ray& ray, Trav& tv) {
  const bool shade = (tv.lcur & LC_SHADE) != 0;
  uint32_t first, count; leaf_range<BVH>(sc, tv, &first, &count);
  const uint32_t i = tv.lcur & 0xFFu; bool have = (tv.lcur & LC_HAVE) != 0;
  const uint32_t idx = shade ? (uint32_t)tv.best_id : first + i;
  WPT_CHECK(idx < sc.num_shapes && (shade || i < count));
  const float4* p = reinterpret_cast<const float4*>(sc.shapes + idx);
  const float4 q0 = __ldg(p), q1 = __ldg(p + 1);
  float t; F3 n = f3(0, 0, 0); bool ent = true;
  const bool hit = torus_trace(q0, q1, ray, &t, shade ? &n : nullptr, &ent);
  if (shade) {
    tv.lf = __float_as_uint(n.x); tv.cnt = __float_as_uint(n.y); tv.sp = __float_as_int(n.z);
    tv.lcur = LC_BEST_TORUS | LC_NREADY | (ent ? LC_ENTERING : 0u);
    return false;
  }
  uint32_t best_torus = tv.lcur & LC_BEST_TORUS;
  if (hit && (have ? (0.0f < t && t < tv.bound) : (t <= tv.bound))) { have = true; tv.bound = t; tv.best_id = (int)(first + i); best_torus = LC_BEST_TORUS; }
  if (i + 1u >= count) { tv.lcur = best_torus; return true; }
  tv.lcur = best_torus | (i + 1u) | (have ? LC_HAVE : 0u) | LC_RESUMED;
  return false;
}
WPT_DEV TorusPre torus_pre(const Trav& tv) {
  TorusPre r; r.n = f3(__uint_as_float(tv.lf), __uint_as_float(tv.cnt), __int_as_float(tv.sp)); r.entering = (tv.lcur & LC_ENTERING) != 0;
  return r;
}
// true if the node the lane is about to enter is a leaf
template <int BVH>
WPT_DEV bool trav_at_leaf(const Trav& tv) { return BVH == 4 ? (int)tv.lf < 0 : tv.cnt != 0; }

WPT_DEV GHit trav_result(const Trav& tv) {
  GHit g;
  g.visits = tv.visits; g.prims = tv.prims;
  // closest (scene.rs:406-422): the BVH hit wins unless the plane hit is strictly closer. Any
  // BVH hit satisfies t <= plane distance, so it wins whenever it exists.
  if (tv.best_id >= 0) { g.t = tv.bound; g.id = tv.best_id; }
  else { g.t = tv.inf_t; g.id = tv.inf_id; }
  return g;
}

template <int BVH, int KIND>
WPT_DEV GHit trace_g_t(const DScene& sc, const Ray& ray) {
  Trav tv;
  uint32_t stack_n[WPT_STACK]; float stack_d[WPT_STACK];
  if (trav_begin<BVH, KIND>(sc, ray, tv)) {
    for (;;) {
      bool need_pop = true;
      if (trav_at_leaf<BVH>(tv)) trav_leaf<BVH, KIND>(sc, ray, tv);
      else need_pop = trav_inner<BVH>(sc, ray, tv, stack_n, stack_d);
      if (need_pop && !trav_pop<BVH>(sc, tv, stack_n, stack_d)) break;
    }
  }
  return trav_result(tv);
}
// generic version (any scene, either BVH): probes, photon emission, the wavefront engine
static __device__ __noinline__ GHit trace_g(const DScene& sc, const Ray& ray) {
  return sc.bvh_kind == 4 ? trace_g_t<4, K_EXT>(sc, ray) : trace_g_t<2, K_EXT>(sc, ray);
}

}  // namespace wpt
