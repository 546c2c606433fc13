// Photon warm-up + octree light-CDF build, adaptive / random sampling drivers and the
// reference-style compute() tick driver. All rendering goes through Context::run_paths.
#include "context.h"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace wpt {

// ------------------------------------------------------------------ photons
// tracer.rs:103-152 in mode B: shots k = 0,1,2,... on their own streams; the photon
// set is every diffuse hit among the shots up to and including the one that produces photon
// number `photon_target`. Shots run in batches on the device; only the last batch needs the
// host to find the cut. Then the octree of photon_tree.rs is built level by level.
void Context::build_photons() {
  require_device();
  if (photons_ready) return;
  const auto t_begin = std::chrono::steady_clock::now();
  const uint32_t L = (uint32_t)scene.lights.size();
  if (L == 0) throw std::runtime_error("Invalid range");   // rng.rs:27 — next_in_range(0, 0) panics
  const uint64_t target = cfg.photon_target;
  const uint32_t batch = 131072;
  const uint32_t cap = (uint32_t)std::min<uint64_t>(target + batch, 0x7FFFFFFFull);
  // Multi-GPU: with a reduce hook rank r emits the shots r, r + world, ... of every
  // batch into dense per-shot slots and the batch is merged with an integer sum-allreduce; every
  // rank then holds the same records and builds the same tree. Without a hook every rank emits all shots.
  const bool split = cfg.world > 1 && (bool)reduce_hook;
  const uint32_t e_rank = split ? cfg.rank : 0u, e_world = split ? cfg.world : 1u;
  ph_loc_w.alloc(cap); ph_light_shot.alloc(cap); ph_dense.alloc((size_t)batch * 6);
  uint32_t* d_meta = ph_dense.p; uint32_t* d_light = ph_dense.p + batch; float4* d_lw = reinterpret_cast<float4*>(ph_dense.p + 2 * (size_t)batch);
  RenderParams rp = params(WPT_NORMAL_NEE);
  // every batch is compacted on the device (stored flags -> exclusive scan -> scatter in shot order, cut at the shot that
  // stores photon number `target`); the host reads back four counters per batch, never the records
  d_seg_cnt.alloc((size_t)batch + 1); d_seg_off.alloc((size_t)batch + 1);
  const size_t sb = seg_scan_bytes(batch);
  d_scan_tmp.alloc(sb ? sb : 1);
  a_stats.alloc(4);
  uint64_t shots = 0, stored = 0, visits = 0, own_shots = 0;
  uint32_t guard = 0;
  ph_host_valid = false; tree_host_valid = false;
  while (stored < target) {
    if (++guard > 100000) throw std::runtime_error("photon warm-up does not converge (no diffuse surface reachable)");
    WPT_CUDA(cudaMemsetAsync(ph_dense.p, 0, (size_t)batch * 6 * sizeof(uint32_t), stream));
    WPT_CUDA(cudaMemsetAsync(a_stats.p, 0, 4 * sizeof(unsigned long long), stream));
    launch_photon_emit(rp, shots, batch, e_rank, e_world, d_meta, d_light, d_lw, stream);
    launches += 1;
    if (split) reduce_hook(ph_dense.p, (uint64_t)batch * 6);
    launch_photon_compact(d_meta, d_light, d_lw, batch, (uint32_t)(target - stored), (uint32_t)stored, e_rank, e_world, d_seg_cnt.p, d_seg_off.p, d_scan_tmp.p, sb,
                          ph_loc_w.p, ph_light_shot.p, a_stats.p, stream);
    launches += 3;
    WPT_CUDA(cudaMemcpyAsync(h_counters + 12, a_stats.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    WPT_CUDA(cudaStreamSynchronize(stream));
    stored += h_counters[12]; shots += h_counters[13]; visits += h_counters[14]; own_shots += h_counters[15];
  }
  photon_shots = shots; photon_count = stored;
  photons_shot_total += shots; photons_stored_total += stored;
  // the photon rays are rays of this session: rays += shots, node visits += their visits
  photon_rays += own_shots; photon_visits += visits;
  // ---- octree topology, level by level (photon_tree.rs:165-196)
  const uint32_t N = (uint32_t)stored;
  std::vector<uint32_t> child_base(1, 0xFFFFFFFFu), count(1, N), depth(1, 0);
  oc_node_of.alloc(N);
  WPT_CUDA(cudaMemsetAsync(oc_node_of.p, 0, N * sizeof(uint32_t), stream));
  size_t frontier_begin = 0;
  for (uint32_t level = 0;; level++) {
    if (level > 48) throw std::runtime_error("photon octree too deep (more than 1024 photons at one point)");
    size_t frontier_end = child_base.size();
    bool any = false;
    for (size_t n = frontier_begin; n < frontier_end; n++)
      if (count[n] > 1024) {   // MAX_PHOTONS_IN_CELL, photon_tree.rs:29
        child_base[n] = (uint32_t)child_base.size();
        for (int c = 0; c < 8; c++) { child_base.push_back(0xFFFFFFFFu); count.push_back(0); depth.push_back(level + 1); }
        any = true;
      }
    if (!any) break;
    oc_child_base.alloc(child_base.size()); oc_count.alloc(child_base.size());
    WPT_CUDA(cudaMemcpyAsync(oc_child_base.p, child_base.data(), child_base.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    WPT_CUDA(cudaMemsetAsync(oc_count.p, 0, child_base.size() * sizeof(uint32_t), stream));
    launch_octree_assign(ph_loc_w.p, N, oc_node_of.p, oc_child_base.p, oc_count.p, stream);
    launches += 1;
    std::vector<uint32_t> cnt(child_base.size());
    WPT_CUDA(cudaMemcpyAsync(cnt.data(), oc_count.p, cnt.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    WPT_CUDA(cudaStreamSynchronize(stream));
    for (size_t n = frontier_end; n < child_base.size(); n++) count[n] = cnt[n];
    frontier_begin = frontier_end;
  }
  const uint32_t nodes = (uint32_t)child_base.size();
  // ---- bins (fixed point) and CDFs
  p_child_base.alloc(nodes); p_cum.alloc((size_t)nodes * L); p_bins.alloc((size_t)nodes * L); oc_fx.alloc((size_t)nodes * L);
  WPT_CUDA(cudaMemcpyAsync(p_child_base.p, child_base.data(), nodes * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  WPT_CUDA(cudaMemsetAsync(oc_fx.p, 0, (size_t)nodes * L * sizeof(unsigned long long), stream));
  launch_octree_bins(ph_loc_w.p, ph_light_shot.p, N, p_child_base.p, oc_fx.p, L, stream);
  launch_octree_cdf(oc_fx.p, p_bins.p, p_cum.p, nodes, L, stream);
  launches += 2;
  p_nodes = nodes;
  // ---- neighbour table for the leaves: the node the reference's walk to the leaf's depth
  // ends in for a point one cell further in each of the 26 directions (same f32 halving as
  // octree_child, photon_tree.rs:235-251)
  {
    struct HB { float lo[3], hi[3]; };
    std::vector<HB> bounds(nodes);
    bounds[0] = HB{{-1024.0f, -1024.0f, -1024.0f}, {1024.0f, 1024.0f, 1024.0f}};
    for (uint32_t n = 0; n < nodes; n++) {   // children always have larger indices than their parent
      if (child_base[n] == 0xFFFFFFFFu) continue;
      const HB& b = bounds[n];
      float c[3] = {0.5f * (b.lo[0] + b.hi[0]), 0.5f * (b.lo[1] + b.hi[1]), 0.5f * (b.lo[2] + b.hi[2])};
      for (int k = 0; k < 8; k++) {
        HB cb;
        int up[3] = {(k >> 2) & 1, (k >> 1) & 1, k & 1};   // octant = 4 [x>=cx] + 2 [y>=cy] + [z>=cz]
        for (int a = 0; a < 3; a++) { cb.lo[a] = up[a] ? c[a] : b.lo[a]; cb.hi[a] = up[a] ? b.hi[a] : c[a]; }
        bounds[child_base[n] + k] = cb;
      }
    }
    auto walk = [&](uint32_t d, const float q[3]) {
      HB b = bounds[0];
      uint32_t nd = 0;
      for (;;) {
        uint32_t cbase = child_base[nd];
        if (cbase == 0xFFFFFFFFu || d == 0) return nd;
        uint32_t idx = 0;
        for (int a = 0; a < 3; a++) {
          float c = 0.5f * (b.lo[a] + b.hi[a]);
          if (q[a] < c) b.hi[a] = c; else { b.lo[a] = c; idx += (a == 0 ? 4u : a == 1 ? 2u : 1u); }
        }
        nd = cbase + idx; d--;
      }
    };
    std::vector<uint32_t> nbr((size_t)nodes * 27, 0u);
    for (uint32_t n = 0; n < nodes; n++) {
      if (child_base[n] != 0xFFFFFFFFu) continue;
      const HB& b = bounds[n];
      for (int k = 0; k < 27; k++) {
        int d3[3] = {k % 3 - 1, (k / 3) % 3 - 1, k / 9 - 1};
        float q[3];
        for (int a = 0; a < 3; a++) {
          float sz = b.hi[a] - b.lo[a];
          float c = 0.5f * (b.lo[a] + b.hi[a]);
          q[a] = c + (float)d3[a] * sz;   // centre of the neighbour cell: exact (dyadic)
        }
        nbr[(size_t)n * 27 + k] = walk(depth[n], q);
      }
    }
    p_nbr.alloc(nbr.size());
    WPT_CUDA(cudaMemcpyAsync(p_nbr.p, nbr.data(), nbr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    WPT_CUDA(cudaStreamSynchronize(stream));
  }
  oc_child_base_h = child_base; oc_count_h = count; oc_depth_h = depth;   // for the lazy read-back copy (photon_tree_host)
  photons_ready = true;
  photon_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  if (std::getenv("WPT_TRACE_PHOTONS")) std::fprintf(stderr, "wpt photon warm-up: %.2f ms (%llu shots, %llu photons, %u octree nodes)\n", photon_build_ms, (unsigned long long)shots, (unsigned long long)stored, nodes);
}

// Read-back copies for the inspection API (wpt_ctx_photon_list / wpt_ctx_photon_tree), made on first use.
void Context::photon_list_host() {
  require_device();
  if (ph_host_valid) return;
  const size_t n = (size_t)photon_count;
  std::vector<float4> lw(n); std::vector<uint2> ls(n);
  if (n) {
    WPT_CUDA(cudaMemcpyAsync(lw.data(), ph_loc_w.p, n * sizeof(float4), cudaMemcpyDeviceToHost, stream));
    WPT_CUDA(cudaMemcpyAsync(ls.data(), ph_light_shot.p, n * sizeof(uint2), cudaMemcpyDeviceToHost, stream));
    WPT_CUDA(cudaStreamSynchronize(stream));
  }
  ph_light.resize(n); ph_loc.resize(n * 3); ph_w.resize(n);
  for (size_t k = 0; k < n; k++) { ph_light[k] = ls[k].x; ph_loc[k * 3] = lw[k].x; ph_loc[k * 3 + 1] = lw[k].y; ph_loc[k * 3 + 2] = lw[k].z; ph_w[k] = lw[k].w; }
  ph_host_valid = true;
}
void Context::photon_tree_host() {
  require_device();
  if (tree_host_valid) return;
  const std::vector<uint32_t>& child_base = oc_child_base_h; const std::vector<uint32_t>& count = oc_count_h; const std::vector<uint32_t>& depth = oc_depth_h;
  const uint32_t nodes = (uint32_t)child_base.size();
  const uint32_t L = (uint32_t)scene.lights.size();
  // DFS pre-order (children in octant order), like the oracle's flattening
  std::vector<float> cum((size_t)nodes * L), bins((size_t)nodes * L);
  WPT_CUDA(cudaMemcpyAsync(cum.data(), p_cum.p, cum.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
  WPT_CUDA(cudaMemcpyAsync(bins.data(), p_bins.p, bins.size() * sizeof(float), cudaMemcpyDeviceToHost, stream));
  WPT_CUDA(cudaStreamSynchronize(stream));
  pt_meta.clear(); pt_cum.clear(); pt_bins.clear();
  std::vector<uint32_t> st(1, 0);
  while (!st.empty()) {
    uint32_t n = st.back(); st.pop_back();
    bool is_node = child_base[n] != 0xFFFFFFFFu;
    pt_meta.push_back(depth[n]); pt_meta.push_back(is_node ? 1u : 0u); pt_meta.push_back(is_node ? 0u : count[n]);
    pt_cum.insert(pt_cum.end(), cum.begin() + (size_t)n * L, cum.begin() + (size_t)(n + 1) * L);
    pt_bins.insert(pt_bins.end(), bins.begin() + (size_t)n * L, bins.begin() + (size_t)(n + 1) * L);
    if (is_node) for (int c = 7; c >= 0; c--) st.push_back(child_base[n] + c);
  }
  tree_host_valid = true;
}

void Context::photon_sample_batch(const float* pts3, const uint32_t* seeds, uint64_t n, uint32_t* light, float* pdf) {
  require_device();
  if (!photons_ready) throw std::runtime_error("photon tree not built");
  if (!n) return;
  DevBuf<float> d_p, d_pdf; DevBuf<uint32_t> d_s, d_l;
  d_p.alloc(n * 3); d_pdf.alloc(n); d_s.alloc(n); d_l.alloc(n);
  WPT_CUDA(cudaMemcpyAsync(d_p.p, pts3, n * 12, cudaMemcpyHostToDevice, stream));
  WPT_CUDA(cudaMemcpyAsync(d_s.p, seeds, n * 4, cudaMemcpyHostToDevice, stream));
  launch_photon_sample_batch(params(WPT_PNEE).photons, d_p.p, d_s.p, n, d_l.p, d_pdf.p, stream);
  WPT_CUDA(cudaMemcpyAsync(light, d_l.p, n * 4, cudaMemcpyDeviceToHost, stream));
  WPT_CUDA(cudaMemcpyAsync(pdf, d_pdf.p, n * 4, cudaMemcpyDeviceToHost, stream));
  WPT_CUDA(cudaStreamSynchronize(stream));
}

// ------------------------------------------------------------------ sampling strategies
Context::Strategy& Context::strategy_for(uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh) {
  for (auto& s : strategies) if (s.rx == rx && s.ry == ry && s.rw == rw && s.rh == rh) return s;
  if (strategies.size() >= 8) strategies.clear();   // regions come and go (split screen, custom regions): bound the buffers kept for them
  strategies.emplace_back();
  Strategy& s = strategies.back();
  s.rx = rx; s.ry = ry; s.rw = rw; s.rh = rh;
  return s;
}
// reset(): every strategy starts over (as a freshly constructed one would), but keeps its device buffers — freeing and
// re-allocating them costs two implicit device synchronisations per buffer and frame
void Context::clear_strategies() {
  for (auto& s : strategies) {
    const uint32_t N = s.rw * s.rh;
    if (s.round_left.n >= N && N) launch_fill_u32(s.round_left.p, N, 0, stream);
    if (s.state.p) WPT_CUDA(cudaMemsetAsync(s.state.p, 0, 16 * sizeof(unsigned long long), stream));
    s.left_total = 0; s.started = false; s.painted = false; s.random_ticks = 0;
  }
}

// error map + {min, avg, max} of the region (sampling_strategy.rs:133-151, mode-B reduction)
void Context::region_error_launch(Strategy& s) {   // error map + {sum, min, max} into a_stats, no host synchronisation
  const uint32_t N = s.rw * s.rh;
  s.mse.alloc(N);
  a_stats.alloc(4);
  static const unsigned long long init[4] = {0ull, 0x7F800000ull, 0ull, 0ull};
  WPT_CUDA(cudaMemcpyAsync(a_stats.p, init, sizeof init, cudaMemcpyHostToDevice, stream));
  launch_error_map(d_accum.p, W, H, s.rx, s.ry, s.rw, s.rh, s.mse.p, a_stats.p, nullptr, stream);
  launches += 1;
}
void Context::region_error(Strategy& s, float stats3[3]) {
  const uint32_t N = s.rw * s.rh;
  region_error_launch(s);
  unsigned long long out[4];
  WPT_CUDA(cudaMemcpyAsync(out, a_stats.p, sizeof out, cudaMemcpyDeviceToHost, stream));
  WPT_CUDA(cudaStreamSynchronize(stream));
  uint32_t mnb = (uint32_t)out[1], mxb = (uint32_t)out[2];
  float mn, mx;
  std::memcpy(&mn, &mnb, 4); std::memcpy(&mx, &mxb, 4);
  stats3[0] = mn;
  stats3[1] = (float)(((double)out[0] * (1.0 / 1099511627776.0)) / (double)N);
  stats3[2] = mx;
}

void Context::error_map(float* mse, float stats3[3]) {
  require_device();
  uint32_t rx, ry, rw, rh;
  region(&rx, &ry, &rw, &rh);
  Strategy& s = strategy_for(rx, ry, rw, rh);
  region_error(s, stats3);
  if (mse) {
    WPT_CUDA(cudaMemcpyAsync(mse, s.mse.p, (size_t)rw * rh * sizeof(float), cudaMemcpyDeviceToHost, stream));
    WPT_CUDA(cudaStreamSynchronize(stream));
  }
}
void Context::round_spp(uint32_t* spp) {
  require_device();
  uint32_t rx, ry, rw, rh;
  region(&rx, &ry, &rw, &rh);
  Strategy& s = strategy_for(rx, ry, rw, rh);
  if (s.round_spp.n < (size_t)rw * rh) throw std::runtime_error("no adaptive / random round has run on this region");
  WPT_CUDA(cudaMemcpyAsync(spp, s.round_spp.p, (size_t)rw * rh * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  WPT_CUDA(cudaStreamSynchronize(stream));
}

// Render `take[]` (region-indexed samples per pixel) on this session's rows. Contract B10: a pixel's samples of the
// call are summed in segments of WPT_SEGMENT_LEN. `bounded` = no pixel has more than 64 samples (adaptive rounds: <= 33),
// so one launch with the device-built segment list does it; otherwise (random strategy) the call is cut into passes of at
// most 64 samples per pixel — the same segments in the same order. The wavefront engine runs one segment per pass.
void Context::render_take(Strategy& s, uint32_t render_type, bool bounded) {
  ensure_slots(s.rx, s.ry, s.rw, s.rh);
  launch_gather_slot_spp(s.take.p, s_pixel.p, slots, W, s.rx, s.ry, s.rw, s_spp.p, stream);
  launches += 1;
  if (!use_wavefront() && bounded) { run_persistent(render_type, s_spp.p, 0); return; }
  const uint32_t per_pass = use_wavefront() ? WPT_SEGMENT_LEN : 64u;
  d_pass_spp.alloc(slots);
  for (uint32_t pass = 0;; pass++) {
    uint32_t any = 0;
    WPT_CUDA(cudaMemsetAsync(w_work.p + 1, 0, sizeof(uint32_t), stream));
    launch_segment_pass_spp(s_spp.p, slots, per_pass, pass, d_pass_spp.p, w_work.p + 1, stream);
    WPT_CUDA(cudaMemcpyAsync(&any, w_work.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    WPT_CUDA(cudaStreamSynchronize(stream));
    launches += 1;
    if (!any) break;
    if (use_wavefront()) run_wavefront(render_type, d_pass_spp.p, 0);
    else run_persistent(render_type, d_pass_spp.p, 0);
  }
}

// AdaptiveSamplingStrategy in mode B (see the oracle's mb_render_adaptive for the contract).
// Every rank evaluates the (cheap) error map of the whole region from the gathered
// accumulators, so no reduction is needed; each rank renders only its own rows.
// AdaptiveSamplingStrategy in mode B (see the oracle's mb_render_adaptive for the contract), device-driven: the strategy's
// state (samples left in the round, ticks used, first-queue flag) lives in s.state and every step — open a round if the last
// one is used up (first queue, or error map -> samples per pixel), cut it to the room left in the budget in the reference's pop
// order, render, subtract — is a fixed sequence of launches whose kernels read that state. The host enqueues steps ahead of
// the device and reads one word (ticks used) every AD_CHUNK steps; steps enqueued past the end of the budget find it spent
// and do nothing. Every rank evaluates the (cheap) error map of the whole region from the gathered accumulators, so no
// reduction is needed; each rank renders only its own rows.
uint64_t Context::run_adaptive(Strategy& s, uint32_t render_type, uint64_t budget, const std::function<void()>& exchange) {
  if (use_wavefront()) return run_adaptive_host(s, render_type, budget, exchange);   // that engine polls the host anyway
  const uint32_t N = s.rw * s.rh;
  if (!N || !budget) return 0;
  if (s.round_left.n < N) {
    s.round_left.alloc(N); s.take.alloc(N); s.round_spp.alloc(N); s.mse.alloc(N); s.state.alloc(16);
    launch_fill_u32(s.round_left.p, N, 0, stream);
    WPT_CUDA(cudaMemsetAsync(s.state.p, 0, 16 * sizeof(unsigned long long), stream));
  }
  const uint32_t blocks = (N + 1023) / 1024;
  a_block_tot.alloc(blocks); a_block_suffix.alloc(blocks);
  unsigned long long* st = s.state.p;
  launch_ad_setup(st, budget, stream);
  static const int chunk = std::getenv("WPT_AD_CHUNK") ? std::max(1, std::atoi(std::getenv("WPT_AD_CHUNK"))) : 4;
  uint64_t used = 0, steps = 0;
  while (used < budget) {
    // steps to enqueue before the next read-back: at most `chunk`, and not more than the budget left is expected to need at the
    // pace of the steps so far (a step past the end of the budget is harmless but costs its launches and, on several GPUs, an all-gather)
    int todo = chunk;
    if (steps && used) todo = (int)std::min<uint64_t>((uint64_t)chunk, std::max<uint64_t>(1, ((budget - used) * steps + used - 1) / used));
    for (int k = 0; k < todo; k++) {
      ev_mark(2, true);
      launch_ad_begin(st, stream);
      launch_ad_first(st, s.round_left.p, s.round_spp.p, N, stream);
      launch_error_map(d_accum.p, W, H, s.rx, s.ry, s.rw, s.rh, s.mse.p, st, st + 9, stream);
      launch_adaptive_spp(s.mse.p, N, st, s.round_left.p, s.round_spp.p, d_sampling.p, W, s.rx, s.ry, s.rw, st + 9, stream);
      launch_ad_total(st, stream);
      launch_cut_device(s.round_left.p, N, a_block_tot.p, a_block_suffix.p, st + 11, s.take.p, stream);
      ev_mark(2, false);
      ev_mark(3, true);
      render_take(s, render_type, true);   // an adaptive round holds at most 33 samples per pixel
      launch_sub_u32(s.round_left.p, s.take.p, N, stream);
      launch_ad_end(st, stream);
      ev_mark(3, false);
      launches += 10;
      if (exchange) { ev_mark(4, true); exchange(); ev_mark(4, false); }   // multi-GPU: gather the other ranks' rows before the next error map
      steps++;
    }
    WPT_CUDA(cudaMemcpyAsync(h_counters + 14, st + 5, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    WPT_CUDA(cudaMemcpyAsync(h_counters + 15, st + 10, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    WPT_CUDA(cudaStreamSynchronize(stream));
    used = h_counters[14];
    if (steps > 4096 + budget / N) throw std::runtime_error("adaptive rounds do not converge");
  }
  adaptive_rounds += h_counters[15];
  return used;
}

// The same strategy with the host in the loop (one read-back per round): the multi-kernel wavefront engine.
uint64_t Context::run_adaptive_host(Strategy& s, uint32_t render_type, uint64_t budget, const std::function<void()>& exchange) {
  // with profiling on, CUDA events split every round into error map / render / exchange (prof_ms[2..4], no extra host syncs)
  static const bool trace = std::getenv("WPT_TRACE_ROUNDS") != nullptr;   // host wall clock per round, to stderr
  auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double w0 = trace ? wall() : 0.0;
  const uint32_t N = s.rw * s.rh;
  if (!N) return 0;
  if (s.round_left.n < N) { s.round_left.alloc(N); s.take.alloc(N); s.round_spp.alloc(N); launch_fill_u32(s.round_left.p, N, 0, stream); s.left_total = 0; s.started = false; }
  const uint32_t blocks = (N + 1023) / 1024;
  a_block_tot.alloc(blocks); a_block_suffix.alloc(blocks);
  uint64_t used = 0;
  while (used < budget) {
    if (s.left_total == 0) {
      if (!s.started) {   // first queue: 4 samples per pixel (sampling_strategy.rs:197-203)
        launch_fill_u32(s.round_left.p, N, 4u, stream);
        s.left_total = 4ull * N;
        s.started = true;
      } else {
        // error map, {sum, min, max} and the samples per pixel stay on the device; the host reads back one word, the
        // round's total (it decides whether the budget ends inside this round)
        ev_mark(2, true);
        region_error_launch(s);
        launch_adaptive_spp(s.mse.p, N, a_stats.p, s.round_left.p, nullptr, d_sampling.p, W, s.rx, s.ry, s.rw, nullptr, stream);
        launches += 1;
        ev_mark(2, false);
        WPT_CUDA(cudaMemcpyAsync(h_counters + 15, a_stats.p + 3, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
        WPT_CUDA(cudaStreamSynchronize(stream));
        s.left_total = h_counters[15];
      }
      WPT_CUDA(cudaMemcpyAsync(s.round_spp.p, s.round_left.p, N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
    }
    uint64_t room = budget - used;
    uint64_t taken;
    if (s.left_total <= room) {
      WPT_CUDA(cudaMemcpyAsync(s.take.p, s.round_left.p, N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
      taken = s.left_total;
    } else {
      // the budget ends inside the round: cut in pop order (from the last pixel backwards)
      launch_cut(s.round_left.p, N, a_block_tot.p, nullptr, 0, nullptr, 0, stream);
      std::vector<unsigned long long> tot(blocks), suf(blocks);
      WPT_CUDA(cudaMemcpyAsync(tot.data(), a_block_tot.p, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      WPT_CUDA(cudaStreamSynchronize(stream));
      unsigned long long run = 0;
      for (uint32_t b = blocks; b-- > 0;) { suf[b] = run; run += tot[b]; }
      WPT_CUDA(cudaMemcpyAsync(a_block_suffix.p, suf.data(), blocks * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream));
      launch_cut(s.round_left.p, N, nullptr, a_block_suffix.p, room, s.take.p, 1, stream);
      launches += 2;
      taken = room;
    }
    const double w1 = trace ? wall() : 0.0;
    ev_mark(3, true);
    render_take(s, render_type, true);   // an adaptive round holds at most 33 samples per pixel
    launch_sub_u32(s.round_left.p, s.take.p, N, stream);
    launches += 1;
    ev_mark(3, false);
    s.left_total -= taken;
    used += taken;
    const double w2 = trace ? wall() : 0.0;
    if (exchange) { ev_mark(4, true); exchange(); ev_mark(4, false); }   // multi-GPU: gather the other ranks' rows before the next error map
    adaptive_rounds += 1;
    if (trace) { const double w3 = wall(); std::fprintf(stderr, "wpt round (rank %u): error map + total %.3f ms, enqueue render %.3f ms, exchange %.3f ms, taken %llu\n", cfg.rank, w1 - w0, w2 - w1, w3 - w2, (unsigned long long)taken); w0 = w3; }
  }
  return used;
}

// RandomSamplingStrategy in mode B: `ticks` pixel picks -> samples per pixel for this call.
void Context::run_random(Strategy& s, uint32_t render_type, uint64_t ticks) {
  const uint32_t N = s.rw * s.rh;
  if (!N || !ticks) return;
  s.take.alloc(N); s.round_spp.alloc(N);
  launch_fill_u32(s.take.p, N, 0, stream);
  uint32_t seed = cfg.base_seed ^ (uint32_t)(s.rx * 0x9E3779B1u + s.ry);
  launch_random_ticks(s.random_ticks, ticks, seed, s.rw, s.rh, s.take.p, stream);
  launches += 2;
  s.random_ticks += ticks;
  WPT_CUDA(cudaMemcpyAsync(s.round_spp.p, s.take.p, N * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
  render_take(s, render_type, ticks <= 64);
}

uint64_t Context::render_adaptive(uint64_t budget) {
  require_device();
  uint32_t rx, ry, rw, rh;
  region(&rx, &ry, &rw, &rh);
  return run_adaptive(strategy_for(rx, ry, rw, rh), cfg.render_type, budget, exchange_hook);
}
void Context::render_random(uint64_t ticks) {
  require_device();
  uint32_t rx, ry, rw, rh;
  region(&rx, &ry, &rw, &rh);
  run_random(strategy_for(rx, ry, rw, rh), cfg.render_type, ticks);
}

// wasm_interface.rs:374-384 + tracer.rs:103-123: n/2 ticks for the left half, the rest for the
// right; a PNEE half first spends ticks on its photon warm-up at 32 shots per tick. Mode-B
// deviation: the warm-up runs to completion in the first call that needs it and is
// charged shots/32 ticks, capped at that call's budget.
void Context::compute(uint64_t num_samples) {
  require_device();
  const uint32_t lw = W / 2;
  struct Half { uint32_t x, w; HalfSettings hs; uint64_t ticks; };
  Half halves[2] = {{0, lw, left, num_samples / 2}, {lw, W - lw, right, num_samples - num_samples / 2}};
  for (const Half& h : halves) {
    if (h.w == 0) continue;
    uint64_t ticks = h.ticks;
    if (h.hs.type == WPT_PNEE && !photons_ready) {
      build_photons();
      uint64_t cost = photon_shots / 32;
      ticks = ticks > cost ? ticks - cost : 0;
    }
    if (!ticks) continue;
    Strategy& s = strategy_for(h.x, 0, h.w, H);
    if (!s.painted) {   // both strategies paint their half blue when created (sampling_strategy.rs:44-50,205-214)
      launch_fill_region_rgba(d_sampling.p, W, h.x, 0, h.w, H, 0xFFFF0000u, stream);
      s.painted = true;
    }
    if (h.hs.adaptive) run_adaptive(s, h.hs.type, ticks, exchange_hook);   // multi-GPU: the error map needs every rank's rows
    else run_random(s, h.hs.type, ticks);
  }
}

}  // namespace wpt
