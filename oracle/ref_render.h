// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED. See ref_core.h.
// Restates src/render_target.rs, src/data/stack.rs, src/graphics/sampling_strategy.rs,
// src/tracer.rs and src/wasm_interface.rs.
//
// Two drivers share the integrator (`trace_original_color`):
//   mode A — the reference's semantics: ONE sequential xorshift32 stream shared by pixel
//            selection, jitter, BSDF sampling, light choice, photons and roulette
//            (rng.rs:10-12, wasm_interface.rs:87), left/right viewport halves.
//   mode B — the contract the GPU implements: per-path / per-shot streams,
//            exact samples per pixel, order-independent photon bins and error sums.
#pragma once
#include "ref_scene.h"
#include "ref_photon.h"
#include <thread>
#include <atomic>
#include <functional>

namespace ref {

static inline Vec3 clamp01(Vec3 v) {   // render_target.rs:184-186
  return Vec3(fmin_(fmax_(v.x, 0.0f), 1.0f), fmin_(fmax_(v.y, 0.0f), 1.0f), fmin_(fmax_(v.z, 0.0f), 1.0f));
}
static inline uint8_t to_u8(float v) {   // `( x.min(1.0).max(0.0) * 255.0 ) as u8` (saturating, NaN -> 0)
  float c = fmax_(fmin_(v, 1.0f), 0.0f) * 255.0f;
  if (!(c > 0.0f)) return 0;
  if (c >= 255.0f) return 255;
  return (uint8_t)c;
}
static const float GAUSS3[9] = {1, 2, 1, 2, 4, 2, 1, 2, 1};
static const float GAUSS5[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};

struct RenderTarget {   // render_target.rs:5-138
  size_t viewport_width, viewport_height;
  std::vector<Vec3> acc_buffer;
  std::vector<uint64_t> acc_count;
  std::vector<uint8_t> result;
  RenderTarget(size_t w, size_t h) : viewport_width(w), viewport_height(h), acc_buffer(w * h), acc_count(w * h, 0), result(w * h * 4, 0) {
    for (size_t i = 0; i < w * h; i++) result[i * 4 + 3] = 255;
  }
  void clear() {
    for (size_t i = 0; i < viewport_width * viewport_height; i++) {
      acc_buffer[i] = Vec3(); acc_count[i] = 0;
      result[i * 4] = result[i * 4 + 1] = result[i * 4 + 2] = 0;
    }
  }
  void write(size_t x, size_t y, Vec3 v) {   // render_target.rs:55-65
    size_t i = viewport_width * y + x;
    acc_buffer[i] = acc_buffer[i] + v;
    acc_count[i] += 1;
    Vec3 a = acc_buffer[i];
    float cnt = (float)acc_count[i];
    result[i * 4 + 0] = to_u8(a.x / cnt);
    result[i * 4 + 1] = to_u8(a.y / cnt);
    result[i * 4 + 2] = to_u8(a.z / cnt);
  }
  // Mode B only (contract B10): `n` samples whose colours were summed from +0 are added at once.
  void write_sum(size_t x, size_t y, Vec3 sum, uint32_t n) {
    size_t i = viewport_width * y + x;
    acc_buffer[i] = acc_buffer[i] + sum;
    acc_count[i] += n;
    Vec3 a = acc_buffer[i];
    float cnt = (float)acc_count[i];
    result[i * 4 + 0] = to_u8(a.x / cnt);
    result[i * 4 + 1] = to_u8(a.y / cnt);
    result[i * 4 + 2] = to_u8(a.z / cnt);
  }
  Vec3 read_clamped(size_t x, size_t y) const {   // render_target.rs:74-77
    size_t i = viewport_width * y + x;
    return clamp01(acc_buffer[i] / (float)acc_count[i]);
  }
  void read_mul(int x, int y, float mul, float* m, Vec3* r) const {   // render_target.rs:132-138
    if (x < 0 || y < 0 || x >= (int)viewport_width || y >= (int)viewport_height) { *m = 0.0f; *r = Vec3(); }
    else { *m = mul; *r = mul * read_clamped((size_t)x, (size_t)y); }
  }
  Vec3 gaussian(size_t x, size_t y, int k, const float* g) const {   // render_target.rs:88-128
    int ix = (int)x, iy = (int)y, h = k / 2;
    float sum = 0.0f; Vec3 acc;
    for (int vy = 0; vy < k; vy++) for (int vx = 0; vx < k; vx++) {
      float m; Vec3 r;
      read_mul(ix + vx - h, iy + vy - h, g[vy * k + vx], &m, &r);
      acc = acc + r;
      sum += m;
    }
    return acc / sum;
  }
  Vec3 gaussian3(size_t x, size_t y) const { return gaussian(x, y, 3, GAUSS3); }
  Vec3 gaussian5(size_t x, size_t y) const { return gaussian(x, y, 5, GAUSS5); }
};

struct SimpleRenderTarget {   // render_target.rs:142-182
  size_t viewport_width, viewport_height;
  std::vector<uint8_t> result;
  SimpleRenderTarget(size_t w, size_t h) : viewport_width(w), viewport_height(h), result(w * h * 4, 0) {
    for (size_t i = 0; i < w * h; i++) result[i * 4 + 3] = 255;
  }
  void clear() { for (size_t i = 0; i < viewport_width * viewport_height; i++) result[i * 4] = result[i * 4 + 1] = result[i * 4 + 2] = 0; }
  void write(size_t x, size_t y, Vec3 v) {
    size_t i = viewport_width * y + x;
    result[i * 4 + 0] = to_u8(v.x); result[i * 4 + 1] = to_u8(v.y); result[i * 4 + 2] = to_u8(v.z);
  }
};

// sampling_strategy.rs:224-230. `a * b * c` on Vec3 * f32 * f32 associates left.
static inline Vec3 mix_color(float v) {
  if (v < 0.5f) return Vec3(0.0f, 1.0f, 0.0f) * (1.0f - 2.0f * v) + Vec3(0.0f, 0.0f, 1.0f) * 2.0f * v;
  return Vec3(0.0f, 0.0f, 1.0f) * (1.0f - 2.0f * (v - 0.5f)) + Vec3(1.0f, 0.0f, 0.0f) * 2.0f * (v - 0.5f);
}
// sampling_strategy.rs:154-163: error -> samples this round, 1..33
static inline float scaled_error(float e, float mn, float avg, float mx) {
  float s = (e < avg) ? 0.5f * ((e - mn) / (avg - mn)) : 0.5f + 0.5f * ((e - avg) / (mx - avg));
  return fmax_(fmin_(s, 1.0f), 0.0f);
}
static inline size_t spp_from_scaled(float s) {
  float c = std::ceil(1.0f + s * 32.0f);
  return c > 0.0f ? (size_t)c : 0;
}

typedef std::pair<uint32_t, uint32_t> Px;

struct SamplingStrategy {
  virtual ~SamplingStrategy() {}
  virtual Px next() = 0;
  virtual void resize(size_t x, size_t y, size_t w, size_t h) = 0;
  virtual void reset() = 0;
};

struct RandomSamplingStrategy : SamplingStrategy {   // sampling_strategy.rs:30-71
  size_t x, y, width, height; Rng* rng;
  RandomSamplingStrategy(size_t x_, size_t y_, size_t w, size_t h, Rng* r, SimpleRenderTarget* st) : x(x_), y(y_), width(w), height(h), rng(r) {
    for (size_t vy = 0; vy < h; vy++) for (size_t vx = 0; vx < w; vx++) st->write(x + vx, y + vy, Vec3(0.0f, 0.0f, 1.0f));
  }
  Px next() override {
    uint32_t px = (uint32_t)(x + rng->next_in_range(0, width));
    uint32_t py = (uint32_t)(y + rng->next_in_range(0, height));
    return Px(px, py);
  }
  void resize(size_t x_, size_t y_, size_t w, size_t h) override { x = x_; y = y_; width = w; height = h; }
  void reset() override {}
};

struct AdaptiveSamplingStrategy : SamplingStrategy {   // sampling_strategy.rs:77-220
  size_t x, y, width, height;
  RenderTarget* target; Rng* rng; SimpleRenderTarget* sampling_target;
  size_t num_sampled = 0;
  std::vector<Px> next_samples;   // data/stack.rs (LIFO)
  AdaptiveSamplingStrategy(size_t x_, size_t y_, size_t w, size_t h, RenderTarget* t, Rng* r, SimpleRenderTarget* st)
      : x(x_), y(y_), width(w), height(h), target(t), rng(r), sampling_target(st) { reset(); }
  Px next() override {
    if (!next_samples.empty()) { Px v = next_samples.back(); next_samples.pop_back(); num_sampled++; return v; }
    std::vector<float> mse(width * height, 0.0f);
    float mse_sum = 0.0f, mse_min = INF_F, mse_max = -INF_F;
    for (size_t yy = 0; yy < height; yy++) for (size_t xx = 0; xx < width; xx++) {
      Vec3 v0 = target->read_clamped(x + xx, y + yy);
      Vec3 v1 = target->gaussian3(x + xx, y + yy);
      Vec3 v2 = target->gaussian5(x + xx, y + yy);
      float e = fmax_(dis_sq(v0, v1), dis_sq(v0, v2));
      mse[yy * width + xx] = e;
      mse_sum += e;
      mse_min = fmin_(mse_min, e);
      mse_max = fmax_(mse_max, e);
    }
    float mse_avg = mse_sum / (float)(width * height);
    for (size_t yy = 0; yy < height; yy++) for (size_t xx = 0; xx < width; xx++) {
      float s = scaled_error(mse[yy * width + xx], mse_min, mse_avg, mse_max);
      size_t spp = spp_from_scaled(s);
      for (size_t k = 0; k < spp; k++) next_samples.push_back(Px((uint32_t)(x + xx), (uint32_t)(y + yy)));
      if (mse_min == mse_max) sampling_target->write(x + xx, y + yy, Vec3());
      else sampling_target->write(x + xx, y + yy, mix_color(s));
    }
    if (next_samples.empty()) throw std::runtime_error("Sampling error");
    Px v = next_samples.back(); next_samples.pop_back();
    return v;   // (the reference does not bump num_sampled here)
  }
  void resize(size_t x_, size_t y_, size_t w, size_t h) override { x = x_; y = y_; width = w; height = h; reset(); }
  void reset() override {   // sampling_strategy.rs:194-219
    next_samples.clear();
    for (size_t vy = 0; vy < height; vy++) for (size_t vx = 0; vx < width; vx++)
      for (int k = 0; k < 4; k++) next_samples.push_back(Px((uint32_t)(x + vx), (uint32_t)(y + vy)));
    for (size_t vy = 0; vy < height; vy++) for (size_t vx = 0; vx < width; vx++) sampling_target->write(x + vx, y + vy, Vec3(0.0f, 0.0f, 1.0f));
    rng->shuffle(next_samples);   // stack.rs:23-28 — same draw pattern as Rng::shuffle
  }
};

struct Camera { Vec3 location; float rot_x, rot_y; };   // tracer.rs:16-26
enum RenderType : int { NoNEE = 0, NormalNEE = 1, PNEE = 2 };   // tracer.rs:28-33, wasm_interface.rs:207-214

struct Stats {
  uint64_t rays = 0, paths = 0, node_visits = 0, photons_shot = 0, photons_stored = 0, prim_tests = 0;
  void add(const Stats& o) { rays += o.rays; paths += o.paths; node_visits += o.node_visits; photons_shot += o.photons_shot; photons_stored += o.photons_stored; prim_tests += o.prim_tests; }
};

enum TrigMode : int { TRIG_LIBM = 0, TRIG_SHARED = 1 };

// material.rs:97-118
static inline void sample_hemisphere(Rng& rng, Vec3 normal, TrigMode trig, Vec3* wi_out, float* pdf_out) {
  float r1 = rng.next();
  float r2 = rng.next();
  float a = 2.0f * PI_F * r1;
  float ca, sa;
  if (trig == TRIG_SHARED) shared_sincos(a, &sa, &ca);
  else { ca = std::cos(a); sa = std::sin(a); }
  float x = ca * std::sqrt(1.0f - r2);
  float y = std::sqrt(r2);
  float z = sa * std::sqrt(1.0f - r2);
  Vec3 x_normal = orthogonal(normal);
  Vec3 z_normal = cross(normal, x_normal);
  Vec3 wi = normalize(x * x_normal + y * normal + z * z_normal);
  *wi_out = wi;
  *pdf_out = dot(wi, normal) / PI_F;
}

// The light chooser: uniform (tracer.rs:275-277) or the photon tree (tracer.rs:271).
struct Integrator {
  const Scene* scene;
  RenderType option;
  bool is_debug_photons;
  TrigMode trig;
  PhotonTree* photons;   // PNEE only

  // tracer.rs:224-330
  Vec3 trace_original_color(const Ray& original_ray, Rng& rng, Stats& st) const {
    bool has_nee = option == NormalNEE || option == PNEE;
    Vec3 color, throughput(1.0f, 1.0f, 1.0f);
    Ray ray = original_ray;
    bool has_diffuse_bounced = false;
    for (;;) {
      bool some; Hit hit;
      st.node_visits += scene->trace(ray, &some, &hit);
      st.rays++;
      if (!some) { color = color + throughput * scene->background.to_vec3(); return color; }
      Vec3 hit_point = ray.at(hit.distance);
      if (hit.mat.emissive) {
        if (is_debug_photons) { if (!has_diffuse_bounced) color = color + throughput * hit.mat.intensity; }
        else if (!has_nee || !has_diffuse_bounced) color = color + throughput * hit.mat.intensity;
        return color;
      }
      // ---- EXTENSION (parity unpinned): specular materials. Not reachable from the reference's scenes.
      if (hit.mat.kind == MAT_REFRACT || (hit.mat.kind == MAT_REFLECT && rng.next() < hit.mat.param)) {
        Vec3 d = ray.dir, n = hit.normal, wi;
        float cos_in = dot(-d, n);
        if (hit.mat.kind == MAT_REFLECT) {
          wi = 2.0f * cos_in * n - (-d);                        // Vec3::reflect, vec3.rs:85-87
          throughput = throughput * hit.mat.color.to_vec3();
        } else {
          float n1 = hit.is_entering ? 1.0f : hit.mat.param, n2 = hit.is_entering ? hit.mat.param : 1.0f;
          if (!hit.is_entering) {                                // Beer's law over the distance travelled inside
            Vec3 a = hit.mat.intensity;
            throughput = throughput * Vec3(shared_exp_neg(a.x * hit.distance), shared_exp_neg(a.y * hit.distance), shared_exp_neg(a.z * hit.distance));
          }
          float r0 = (n1 - n2) / (n1 + n2); r0 = r0 * r0;        // Schlick with total internal reflection
          float cosx = cos_in; bool tir = false;
          if (n1 > n2) { float nr = n1 / n2; float sin2 = nr * nr * (1.0f - cosx * cosx); if (sin2 > 1.0f) tir = true; else cosx = std::sqrt(1.0f - sin2); }
          float x = 1.0f - cosx;
          float fres = tir ? 1.0f : r0 + (1.0f - r0) * x * x * x * x * x;
          if (rng.next() < fres) wi = 2.0f * cos_in * n - (-d);
          else {
            float eta = n1 / n2;
            float k = 1.0f - eta * eta * (1.0f - cos_in * cos_in);
            wi = normalize(eta * d + (eta * cos_in - std::sqrt(fmax_(k, 0.0f))) * n);
          }
        }
        ray = Ray(hit_point + wi * EPSILON, wi);
        has_diffuse_bounced = false;                            // a light seen through a specular bounce is not covered by NEE
        float keep = fmax_(fmin_(fmax_(fmax_(throughput.x, throughput.y), throughput.z), 0.9f), 0.1f);
        if (rng.next() < keep) { throughput = throughput * (1.0f / keep); continue; }
        return color;
      }
      Vec3 wi; float pdf;
      sample_hemisphere(rng, hit.normal, trig, &wi, &pdf);
      Color3 brdf = hit.mat.color / PI_F;                       // material.rs:120-126
      float cos_i = dot(wi, hit.normal);
      throughput = throughput * brdf.to_vec3() * cos_i / pdf;   // tracer.rs:262
      ray = Ray(hit_point + wi * EPSILON, wi);
      has_diffuse_bounced = true;
      if (has_nee) {
        size_t light_id; float light_chance;
        if (option == PNEE) photons->sample(rng, hit_point, &light_id, &light_chance);
        else { size_t nl = scene->lights.size(); light_id = rng.next_in_range(0, nl); light_chance = 1.0f / (float)nl; }
        size_t light_shape_id = scene->lights[light_id];
        const Shape& light_shape = scene->shapes[light_shape_id];
        Vec3 point_on_light, light_normal, intensity;
        light_shape.pick_random(rng, &point_on_light, &light_normal, &intensity);
        Vec3 to_light = point_on_light - hit_point;
        float dsq = len_sq(to_light);
        to_light = to_light / std::sqrt(dsq);
        float cos_i2 = dot(to_light, hit.normal);
        float cos_o = dot(-to_light, light_normal);
        if (cos_i2 > 0.0f && cos_o > 0.0f) {
          if (is_debug_photons) color = color + throughput * intensity;
          else {
            bool occluded;
            st.node_visits += scene->shadow_ray(hit_point, point_on_light, light_shape_id, &occluded);
            st.rays++;
            if (!occluded) {
              float solid_angle = (light_shape.surface_area() * cos_o) / dsq;
              color = color + throughput * intensity * solid_angle * cos_i2 * (1.0f / light_chance);
            }
          }
        }
      }
      float keep_chance = fmax_(fmin_(fmax_(fmax_(throughput.x, throughput.y), throughput.z), 0.9f), 0.1f);   // tracer.rs:318
      if (rng.next() < keep_chance) throughput = throughput * (1.0f / keep_chance);
      else return color;
    }
  }

  // tracer.rs:131-147 — one photon shot. Returns true if a photon was produced.
  bool shoot_photon(Rng& rng, Stats& st, size_t* light_out, Vec3* loc, float* w) const {
    size_t light_id = rng.next_in_range(0, scene->lights.size());
    const Shape& light_shape = scene->shapes[scene->lights[light_id]];
    Vec3 pol, ln, intensity;
    light_shape.pick_random(rng, &pol, &ln, &intensity);
    Vec3 light_normal = rng.next_hemisphere(ln);
    Ray ray(pol + light_normal * EPSILON, light_normal);
    bool some; Hit hit;
    st.node_visits += scene->trace(ray, &some, &hit);
    st.rays++;
    st.photons_shot++;
    if (some) {
      Vec3 hp = ray.at(hit.distance) + hit.normal * EPSILON;
      if (hit.mat.is_diffuse()) {
        *light_out = light_id; *loc = hp;
        *w = dot(ln, light_normal) * fmax_(fmax_(intensity.x, intensity.y), intensity.z);
        st.photons_stored++;
        return true;
      }
    }
    return false;
  }
};

// tracer.rs:176-191 — primary ray through pixel (x,y) with jitter (j1,j2)
static inline Ray camera_ray(const Camera& cam, size_t W, size_t H, size_t x, size_t y, float j1, float j2) {
  float fw = (float)W, fh = (float)H;
  float w_inv = 1.0f / fw, h_inv = 1.0f / fh, ar = fw / fh;
  float fx = (((float)x + j1) * w_inv - 0.5f) * ar;
  float fy = 0.5f - ((float)y + j2) * h_inv;
  Vec3 pixel(fx, fy, 0.8f);
  Vec3 dir = rot_y(rot_x(normalize(pixel), cam.rot_x), cam.rot_y);
  return Ray(cam.location, dir);
}

static const size_t TOTAL_PHOTONS_NEEDED = 300000;   // tracer.rs:104

// ================================================================ mode A
struct RenderInstance {   // tracer.rs:35-123
  RenderType option;
  Camera* camera; const Scene* scene; Rng* rng; RenderTarget* target;
  std::unique_ptr<SamplingStrategy> strategy;
  bool is_debug_photons;
  TrigMode trig = TRIG_LIBM;
  PhotonTree photons;
  size_t num_photons = 0;
  Stats stats;
  RenderInstance(const Scene* sc, Camera* cam, Rng* r, std::unique_ptr<SamplingStrategy> strat, bool dbg, RenderTarget* t, RenderType opt)
      : option(opt), camera(cam), scene(sc), rng(r), target(t), strategy(std::move(strat)), is_debug_photons(dbg), photons(sc->lights.size(), false) { reset(); }
  void reset() { stats = Stats(); strategy->reset(); }
  void resize(size_t x, size_t y, size_t w, size_t h) { strategy->resize(x, y, w, h); reset(); }
  void update_scene(const Scene* sc) { num_photons = 0; photons = PhotonTree(sc->lights.size(), false); scene = sc; reset(); }
  Integrator integ() { return Integrator{scene, option, is_debug_photons, trig, &photons}; }
  void preprocess_photons(size_t n) {   // tracer.rs:126-152
    Integrator I = integ();
    for (size_t i = 0; i < n; i++) {
      size_t l; Vec3 loc; float w;
      if (I.shoot_photon(*rng, stats, &l, &loc, &w)) { photons.insert(l, loc, w); num_photons++; }
    }
  }
  void compute_rays(size_t n) {   // tracer.rs:156-201
    Integrator I = integ();
    for (size_t i = 0; i < n; i++) {
      Px p = strategy->next();
      float j1 = rng->next();
      float j2 = rng->next();
      Ray ray = camera_ray(*camera, target->viewport_width, target->viewport_height, p.first, p.second, j1, j2);
      Vec3 res = I.trace_original_color(ray, *rng, stats);
      stats.paths++;
      target->write(p.first, p.second, res);
    }
  }
  void compute(size_t num_ticks) {   // tracer.rs:103-123
    if (option == PNEE && num_photons < TOTAL_PHOTONS_NEEDED) {
      size_t n = std::min(TOTAL_PHOTONS_NEEDED - num_photons, num_ticks * 32);
      preprocess_photons(n);
      size_t ticks_left = num_ticks - n / 32;
      while (ticks_left > 0 && num_photons < TOTAL_PHOTONS_NEEDED) {
        size_t m = std::min(TOTAL_PHOTONS_NEEDED - num_photons, ticks_left * 32);
        preprocess_photons(m);
        ticks_left -= m / 32;
      }
      compute_rays(ticks_left);
    } else compute_rays(num_ticks);
  }
};

}  // namespace ref
