// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED. See ref_core.h.
// Session = src/wasm_interface.rs restated (mode A), plus the mode-B driver that
// defines what the CUDA product must reproduce bit for bit (DESIGN.md "mode B").
#pragma once
#include "ref_render.h"
#include <map>

namespace ref {

struct ModeBSettings {
  RenderType type = NormalNEE;
  bool bvh4 = false;
  bool light_debug = false;
  TrigMode trig = TRIG_SHARED;
  uint32_t base_seed = 0xBABABEBEu;
  size_t photon_target = TOTAL_PHOTONS_NEEDED;
  // logical sampling region inside the viewport (full frame by default)
  size_t rx = 0, ry = 0, rw = 0, rh = 0;
};

struct Session {
  // ## global state (wasm_interface.rs:38-42)
  std::map<uint32_t, std::vector<float>> mesh_preload;   // Mesh::Preload
  std::map<uint32_t, std::vector<Shape>> mesh_tris;      // Mesh::Triangled
  std::map<uint32_t, std::vector<uint8_t>> textures;     // stored, never sampled (F5) — except by the extension scene 256
  std::map<uint32_t, std::pair<uint32_t, uint32_t>> tex_size;
  std::unique_ptr<Texture> ext_tex0;
  Rng rng;
  // ## session state
  size_t W, H;
  std::unique_ptr<RenderTarget> target;
  std::unique_ptr<SimpleRenderTarget> sampling_target;
  uint32_t scene_id;
  std::unique_ptr<Scene> scene;
  Camera camera;
  std::unique_ptr<RenderInstance> left, right;   // mode A
  bool use_bvh4 = false;                         // not reachable in the reference (F6)
  TrigMode trig_a = TRIG_LIBM;

  // ## mode B state
  ModeBSettings mb;
  std::unique_ptr<PhotonTree> mb_photons;
  std::vector<PhotonRec> mb_photon_list;
  uint64_t mb_shots = 0;
  Stats mb_stats;
  std::vector<uint32_t> mb_round_spp;   // last adaptive round allocation (debug / tests)

  // wasm_interface.rs:389-398
  std::unique_ptr<Scene> select_scene(uint32_t id) {
    std::unique_ptr<Scene> s(new Scene());
    if (id == 0) s->build(Color3(0, 0, 0), museum_shapes(), use_bvh4);
    else if (id == 2) {
      auto it = mesh_tris.find(1);   // MESH_BUNNY_HIGH, scenes.rs:12
      s->build(Color3(0, 0, 0), bunny_shapes(it == mesh_tris.end() ? nullptr : &it->second), use_bvh4);
    } else if (id == 256) {   // EXTENSION scene
      auto it = textures.find(0);
      ext_tex0.reset();
      if (it != textures.end()) { ext_tex0.reset(new Texture()); ext_tex0->width = tex_size[0].first; ext_tex0->height = tex_size[0].second; ext_tex0->data.assign(it->second.begin(), it->second.begin() + (size_t)ext_tex0->width * ext_tex0->height * 3); }
      s->build(Color3(135.0f / 255.0f, 206.0f / 255.0f, 250.0f / 255.0f), whitted_shapes(ext_tex0.get()), use_bvh4);
    } else throw std::runtime_error("Invalid scene");
    return s;
  }

  // wasm_interface.rs:67-113
  Session(uint32_t w, uint32_t h, uint32_t sid, float cx, float cy, float cz, float rx, float ry, bool bvh4 = false) : W(w), H(h), scene_id(sid), use_bvh4(bvh4) {
    camera = Camera{Vec3(cx, cy, cz), rx, ry};
    target.reset(new RenderTarget(w, h));
    sampling_target.reset(new SimpleRenderTarget(w, h));
    scene = select_scene(sid);
    size_t lw = w / 2;
    std::unique_ptr<SamplingStrategy> ls(new RandomSamplingStrategy(0, 0, lw, h, &rng, sampling_target.get()));
    std::unique_ptr<SamplingStrategy> rs(new AdaptiveSamplingStrategy(lw, 0, w - lw, h, target.get(), &rng, sampling_target.get()));
    left.reset(new RenderInstance(scene.get(), &camera, &rng, std::move(ls), false, target.get(), NormalNEE));
    right.reset(new RenderInstance(scene.get(), &camera, &rng, std::move(rs), false, target.get(), PNEE));
    mb.rw = w; mb.rh = h;
  }
  void reset() {   // wasm_interface.rs:137-148
    target->clear(); sampling_target->clear(); left->reset(); right->reset();
    mb_stats = Stats(); mb_round_left.clear(); mb_adaptive_started = false; mb_random_ticks = 0;
  }
  void update_scene(uint32_t sid) {   // wasm_interface.rs:154-168
    scene_id = sid;
    std::unique_ptr<Scene> ns = select_scene(sid);
    target->clear(); sampling_target->clear();
    left->update_scene(ns.get()); right->update_scene(ns.get());
    scene = std::move(ns);
    mb_photons.reset(); mb_photon_list.clear(); mb_shots = 0; mb_stats = Stats();
    mb_round_left.clear(); mb_adaptive_started = false; mb_random_ticks = 0;
  }
  void update_settings(uint32_t lt, uint32_t rt, uint32_t la, uint32_t ra, uint32_t dbg) {   // wasm_interface.rs:173-214
    if (lt > 2 || rt > 2) throw std::runtime_error("Invalid RenderType magic number");
    size_t lw = W / 2;
    std::unique_ptr<SamplingStrategy> ls, rs;
    if (la == 1) ls.reset(new AdaptiveSamplingStrategy(0, 0, lw, H, target.get(), &rng, sampling_target.get()));
    else ls.reset(new RandomSamplingStrategy(0, 0, lw, H, &rng, sampling_target.get()));
    if (ra == 1) rs.reset(new AdaptiveSamplingStrategy(lw, 0, W - lw, H, target.get(), &rng, sampling_target.get()));
    else rs.reset(new RandomSamplingStrategy(lw, 0, W - lw, H, &rng, sampling_target.get()));
    target->clear(); sampling_target->clear();
    left.reset(new RenderInstance(scene.get(), &camera, &rng, std::move(ls), dbg == 1, target.get(), (RenderType)lt));
    right.reset(new RenderInstance(scene.get(), &camera, &rng, std::move(rs), dbg == 1, target.get(), (RenderType)rt));
    left->trig = right->trig = trig_a;
  }
  void update_viewport(uint32_t w, uint32_t h) {   // wasm_interface.rs:219-232
    W = w; H = h;
    target.reset(new RenderTarget(w, h));
    sampling_target.reset(new SimpleRenderTarget(w, h));
    // the strategies hold raw pointers to the targets: re-point them
    auto repoint = [&](RenderInstance* ri) {
      ri->target = target.get();
      if (auto* a = dynamic_cast<AdaptiveSamplingStrategy*>(ri->strategy.get())) { a->target = target.get(); a->sampling_target = sampling_target.get(); }
    };
    repoint(left.get()); repoint(right.get());
    size_t lw = w / 2;
    left->resize(0, 0, lw, h);
    right->resize(lw, 0, w - lw, h);
    reset();
    mb.rx = mb.ry = 0; mb.rw = w; mb.rh = h;
  }
  void update_camera(float x, float y, float z, float rx, float ry) {   // wasm_interface.rs:239-248
    camera = Camera{Vec3(x, y, z), rx, ry};
    reset();
  }
  void allocate_mesh(uint32_t id, uint32_t nv) { mesh_tris.erase(id); mesh_preload[id] = std::vector<float>((size_t)nv * 3, 0.0f); }
  float* mesh_vertices(uint32_t id) {
    auto it = mesh_preload.find(id);
    if (it == mesh_preload.end()) throw std::runtime_error("Mesh not allocated");
    return it->second.data();
  }
  bool notify_mesh_loaded(uint32_t id) {   // wasm_interface.rs:293-329
    auto it = mesh_preload.find(id);
    if (it != mesh_preload.end()) {
      mesh_tris[id] = mesh_to_triangles(it->second.data(), it->second.size() / 3);
      mesh_preload.erase(it);
    }
    if ((id == 0 && scene_id == 1) || (id == 1 && scene_id == 2) || (id == 2 && scene_id == 3)) { update_scene(scene_id); return true; }
    return false;
  }
  void compute(size_t n) {   // wasm_interface.rs:374-384
    size_t nl = n / 2;
    left->compute(nl);
    right->compute(n - nl);
  }
  void rebuild_bvh(bool bvh4) { use_bvh4 = bvh4; update_scene(scene_id); }

  // ================================================================ mode B
  Integrator mb_integ() { return Integrator{scene.get(), mb.type, mb.light_debug, mb.trig, mb_photons.get()}; }

  static void finalize_cdfs(Octree& o) { o.cdf.recheck_cdf(); for (auto& c : o.children) finalize_cdfs(c); }

  // Photon warm-up: shots k = 0,1,2,... each on stream (k, 0, STREAM_PHOTON); the photon
  // set is every diffuse hit among the shots up to and including the one that produces
  // photon number `photon_target` (what tracer.rs:103-117 converges to).
  void mb_build_photons(unsigned threads = 1) {
    if (mb_photons) return;
    Integrator I = mb_integ();
    mb_photon_list.clear();
    mb_shots = 0;
    const size_t chunk = 65536;
    struct Rec { bool ok; PhotonRec p; };
    std::vector<Rec> buf(chunk);
    while (mb_photon_list.size() < mb.photon_target) {
      uint64_t k0 = mb_shots;
      std::vector<Stats> tst(threads);
      auto work = [&](unsigned t) {
        for (size_t i = t; i < chunk; i += threads) {
          Rng r(stream_seed((uint32_t)(k0 + i), 0, STREAM_PHOTON, mb.base_seed));
          size_t l; Vec3 loc; float w;
          buf[i].ok = I.shoot_photon(r, tst[t], &l, &loc, &w);
          if (buf[i].ok) buf[i].p = PhotonRec{l, loc, w};
        }
      };
      run_threads(threads, work);
      // count only the shots up to the cut
      size_t used = chunk;
      for (size_t i = 0; i < chunk; i++) {
        if (buf[i].ok) { mb_photon_list.push_back(buf[i].p); if (mb_photon_list.size() == mb.photon_target) { used = i + 1; break; } }
      }
      mb_shots += used;
      if (used == chunk) for (auto& s : tst) mb_stats.add(s);
      else {   // recount the partial chunk exactly
        Stats s;
        for (size_t i = 0; i < used; i++) { Rng r(stream_seed((uint32_t)(k0 + i), 0, STREAM_PHOTON, mb.base_seed)); size_t l; Vec3 loc; float w; I.shoot_photon(r, s, &l, &loc, &w); }
        mb_stats.add(s);
      }
    }
    mb_photons.reset(new PhotonTree(scene->lights.size(), true));
    for (auto& p : mb_photon_list) mb_photons->insert(p.light, p.loc, p.w);
    finalize_cdfs(mb_photons->root);
  }

  static void run_threads(unsigned n, const std::function<void(unsigned)>& f) {
    if (n <= 1) { f(0); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < n; t++) th.emplace_back(f, t);
    for (auto& t : th) t.join();
  }

  // One sample `s` of pixel (x,y): stream (pixel index in the full viewport, s, STREAM_PATH).
  Vec3 mb_sample(const Integrator& I, size_t x, size_t y, uint32_t s, Stats& st) {
    Rng r(stream_seed((uint32_t)(y * W + x), s, STREAM_PATH, mb.base_seed));
    float j1 = r.next();
    float j2 = r.next();
    Ray ray = camera_ray(camera, W, H, x, y, j1, j2);
    Vec3 res = I.trace_original_color(ray, r, st);
    st.paths++;
    return res;
  }
  // Mode-B accumulation contract: the `n` samples a pixel receives in one call are
  // summed in segments of at most `seg` consecutive samples, each segment from +0 in sample order,
  // and the segment sums are added to the accumulator in order. (A single call of <= seg samples
  // on a cleared pixel gives the same bits as RenderTarget::write per sample, since 0 + c == c.)
  // Segments are what lets the GPU run the samples of ONE pixel on several lanes at once.
  static constexpr uint32_t MB_SEGMENT = 8;   // measured: 8 beats 16 by 6 % on the 16-spp bench frame (gpurun_out/sweep9.log)
  void mb_samples(const Integrator& I, size_t x, size_t y, uint32_t n, uint32_t seg, Stats& st) {
    uint32_t s0 = (uint32_t)target->acc_count[y * W + x];
    for (uint32_t j = 0; j < n; j += seg) {
      uint32_t m = n - j < seg ? n - j : seg;
      Vec3 sum(0.0f, 0.0f, 0.0f);
      for (uint32_t k = 0; k < m; k++) sum = sum + mb_sample(I, x, y, s0 + j + k, st);
      target->write_sum(x, y, sum, m);
    }
  }

  // `spp` more samples for every pixel of the region; pixel rows interleaved over threads.
  void mb_render_exact(uint32_t spp, unsigned threads = 1) {
    if (mb.type == PNEE) mb_build_photons(threads);
    Integrator I = mb_integ();
    std::vector<Stats> tst(threads);
    auto work = [&](unsigned t) {
      tl_prim_tests() = 0;
      for (size_t yy = t; yy < mb.rh; yy += threads)
        for (size_t xx = 0; xx < mb.rw; xx++) {
          size_t x = mb.rx + xx, y = mb.ry + yy;
          mb_samples(I, x, y, spp, MB_SEGMENT, tst[t]);
        }
      tst[t].prim_tests = tl_prim_tests();
    };
    run_threads(threads, work);
    for (auto& s : tst) mb_stats.add(s);
  }

  // Error map of the region (sampling_strategy.rs:133-149) with the mode-B reduction:
  // sum in 2^-40 fixed point (order independent), min, max.
  void mb_error_stats(std::vector<float>& mse, float* mn, float* avg, float* mx) {
    mse.assign(mb.rw * mb.rh, 0.0f);
    uint64_t sum_fx = 0; float e_min = INF_F, e_max = -INF_F;
    for (size_t yy = 0; yy < mb.rh; yy++) for (size_t xx = 0; xx < mb.rw; xx++) {
      size_t x = mb.rx + xx, y = mb.ry + yy;
      Vec3 v0 = target->read_clamped(x, y), v1 = target->gaussian3(x, y), v2 = target->gaussian5(x, y);
      float e = fmax_(dis_sq(v0, v1), dis_sq(v0, v2));
      mse[yy * mb.rw + xx] = e;
      sum_fx += photon_weight_fx(e);
      e_min = fmin_(e_min, e); e_max = fmax_(e_max, e);
    }
    *mn = e_min; *mx = e_max;
    *avg = (float)(((double)sum_fx * (1.0 / PHOTON_FX_SCALE)) / (double)(mb.rw * mb.rh));
  }

  // Adaptive sampling with a tick budget (sampling_strategy.rs:77-220 in mode B).
  // The queue of the reference becomes `mb_round_left[pixel]` = samples of the current round
  // still to be taken. A new round starts when the queue is empty: the first round after a
  // reset queues 4 samples per pixel (sampling_strategy.rs:194-203), later rounds 1..33 from
  // the error map (:133-166). Ticks are consumed in the reference's pop order for rounds —
  // LIFO over raster-order pushes, i.e. from the last pixel of the region backwards, each
  // pixel's entries contiguous — so a budget that ends inside a round leaves the rest of the
  // queue for the next call, exactly like compute(n) popping n entries. (The reference
  // shuffles only the very first queue; with per-path streams the order inside a round does
  // not change any sample, only which pixels a partial budget reaches.) Returns ticks used.
  std::vector<uint32_t> mb_round_left;
  bool mb_adaptive_started = false;
  uint64_t mb_render_adaptive(uint64_t budget, unsigned threads = 1) {
    if (mb.type == PNEE) mb_build_photons(threads);
    Integrator I = mb_integ();
    uint64_t used = 0;
    size_t N = mb.rw * mb.rh;
    if (mb_round_left.size() != N) { mb_round_left.assign(N, 0); mb_adaptive_started = false; }
    std::vector<uint32_t> take(N);
    while (used < budget) {
      bool empty = true;
      for (size_t i = 0; i < N && empty; i++) if (mb_round_left[i]) empty = false;
      if (empty) {
        if (!mb_adaptive_started) { std::fill(mb_round_left.begin(), mb_round_left.end(), 4u); mb_adaptive_started = true; }
        else {
          std::vector<float> mse; float mn, avg, mx;
          mb_error_stats(mse, &mn, &avg, &mx);
          for (size_t i = 0; i < N; i++) {
            float sc = scaled_error(mse[i], mn, avg, mx);
            mb_round_left[i] = (uint32_t)spp_from_scaled(sc);
            size_t x = mb.rx + i % mb.rw, y = mb.ry + i / mb.rw;
            if (mn == mx) sampling_target->write(x, y, Vec3()); else sampling_target->write(x, y, mix_color(sc));
          }
        }
        mb_round_spp = mb_round_left;
      }
      // budget cut, from the last pixel backwards
      uint64_t left_ticks = budget - used;
      for (size_t i = N; i-- > 0;) {
        take[i] = (uint64_t)mb_round_left[i] > left_ticks ? (uint32_t)left_ticks : mb_round_left[i];
        left_ticks -= take[i];
      }
      std::vector<Stats> tst(threads);
      auto work = [&](unsigned t) {
        tl_prim_tests() = 0;
        for (size_t yy = t; yy < mb.rh; yy += threads)
          for (size_t xx = 0; xx < mb.rw; xx++) {
            size_t x = mb.rx + xx, y = mb.ry + yy;
            uint32_t n = take[yy * mb.rw + xx];
            if (n) mb_samples(I, x, y, n, MB_SEGMENT, tst[t]);
          }
        tst[t].prim_tests = tl_prim_tests();
      };
      run_threads(threads, work);
      for (auto& s : tst) mb_stats.add(s);
      for (size_t i = 0; i < N; i++) { used += take[i]; mb_round_left[i] -= take[i]; }
    }
    return used;
  }

  // Random strategy in mode B (sampling_strategy.rs:56-59): tick t of the region picks its
  // pixel from stream (t, 0, STREAM_PIXEL) with the reference's two next_in_range draws; the
  // ticks of one call only decide HOW MANY samples each pixel receives, the samples themselves
  // are the pixel's next sample indices (per-path streams), accumulated in index order.
  uint64_t mb_random_ticks = 0;
  void mb_render_random(uint64_t ticks, unsigned threads = 1) {
    if (mb.type == PNEE) mb_build_photons(threads);
    Integrator I = mb_integ();
    size_t N = mb.rw * mb.rh;
    std::vector<uint32_t> take(N, 0);
    for (uint64_t k = 0; k < ticks; k++) {
      Rng r(stream_seed((uint32_t)(mb_random_ticks + k), (uint32_t)((mb_random_ticks + k) >> 32), STREAM_PIXEL, mb.base_seed ^ (uint32_t)(mb.rx * 0x9E3779B1u + mb.ry)));
      size_t x = r.next_in_range(0, mb.rw);
      size_t y = r.next_in_range(0, mb.rh);
      take[y * mb.rw + x]++;
    }
    mb_random_ticks += ticks;
    mb_round_spp = take;
    std::vector<Stats> tst(threads);
    auto work = [&](unsigned t) {
      for (size_t yy = t; yy < mb.rh; yy += threads)
        for (size_t xx = 0; xx < mb.rw; xx++) {
          size_t x = mb.rx + xx, y = mb.ry + yy;
          uint32_t n = take[yy * mb.rw + xx];
          if (n) mb_samples(I, x, y, n, MB_SEGMENT, tst[t]);
        }
    };
    run_threads(threads, work);
    for (auto& s : tst) mb_stats.add(s);
  }

  // Primary-ray probe for the bit-exact gate: sample 0 of every pixel of the viewport.
  void mb_primary_probe(int32_t* ids, uint32_t* visits, float* dist) {
    for (size_t y = 0; y < H; y++) for (size_t x = 0; x < W; x++) {
      Rng r(stream_seed((uint32_t)(y * W + x), 0, STREAM_PATH, mb.base_seed));
      float j1 = r.next();
      float j2 = r.next();
      Ray ray = camera_ray(camera, W, H, x, y, j1, j2);
      GHit g; size_t d = scene->trace_g(ray, &g);
      size_t i = y * W + x;
      if (ids) ids[i] = g.some ? (int32_t)g.shape : -1;
      if (visits) visits[i] = (uint32_t)d;
      if (dist) dist[i] = g.some ? g.dis : INF_F;
    }
  }
};

}  // namespace ref
