// k_wpool — the warp-pool persistent path kernel (engine 0).
//
// The same path tracer as k_mega (one lane = one path at a time, path regeneration, contract B10), but a warp owns a
// POOL of path contexts (P.pool_ctx of them, 160 B each, in global memory — L2 resident) instead of one per lane, and
// strictly alternates between two loops that never share registers:
//
//   logic run  — every lane holds one context in registers and runs the two-half logic pass of k_mega (half A: resolve
//                a shadow-ray result or generate the next camera ray, then the plane tests + root guard of the new ray;
//                half B: shade the hit of a camera / bounce ray, then the same for the ray it produces). A ray that ends
//                at the root guard (85 % of the bunny scene's rays) is consumed by the same lane in the next half; a ray
//                that has to enter the BVH PARKS its context (9 x 16 B stores) in the warp's traversal queue and the
//                lane takes another context — one whose traversal has finished (logic queue), or a fresh segment.
//   trav burst — starts when a full warp of rays waits: every lane pops a job (ray + bound, 3 x 16 B loads) and runs
//                k_mega's voted inner / leaf / pop steps; a lane whose ray is done writes (t, id) into the context,
//                pushes it to the logic queue and refills itself from the traversal queue. When the queue is empty and
//                few lanes are left, their jobs are suspended (6 words into the context; the stack stays in the lane's
//                local memory) and the warp returns to logic.
//
// So both loops run with (almost) all 32 lanes — k_mega, with one path pinned per lane, keeps about half of the lanes
// waiting in the other stage (profiles/r1c_k_mega_bench_full.md: 14.7 lanes in the shading code, 13.9 in the inner-node
// step). The queues are per warp and warp-synchronous (ballot + popc, head / count in uniform registers): no atomics,
// no fences, no spinning, nothing shared between warps but the slot counter. A context's samples are still traced one
// after the other and summed in sample order, so every bit of the result is unchanged (the engines are compared bit for
// bit in tests/test_gpu_features.py).
#include "kernels.h"
#include "device_shade.cuh"

namespace wpt {

#define WP_THREADS 128
#define WP_WARPS (WP_THREADS / 32)
#define WP_QCAP 128u      // queue capacity per warp = largest pool (context ids are bytes)
#define WP_CTX_F4 10u     // 16-byte words per context

// Context record (10 x float4):
//   q0 = ray origin, w = result t / traversal bound        q1 = ray direction, w = bits(result id / plane-hit id)
//   q2 = throughput, w = bits(rng)                          q3 = colour, w = bits(what | bounced << 2 | alive_after_shadow << 3)
//   q4 = segment sum, w = bits(px | py << 16)               q5 = bits(s, s_end, slot), sh_len
//   q6 = bounce origin, w = bits(light shape)               q7 = bounce direction      q8 = shadow contribution   (what == ST_SHADOW only)
//   q9 = traversal state of a queued / suspended job: bits(lf, cnt, best_id, sp)
enum : int { WS_GEN = 0, WS_EXTEND = 1, WS_SHADOW = 2 };

struct WCur {
  F3 o, d; float res_t; int res_id;
  PathRegs ps; int what; bool alive;
  F3 acc; uint32_t pixp, s, s_end, slot;
  float sh_len; int sh_light; F3 ext_o, ext_d, contrib;
};

WPT_DEV void ctx_store(float4* c, const WCur& k, bool trav, uint32_t root_lf, uint32_t root_cnt) {
  c[0] = make_float4(k.o.x, k.o.y, k.o.z, k.res_t);
  c[1] = make_float4(k.d.x, k.d.y, k.d.z, __int_as_float(k.res_id));
  c[2] = make_float4(k.ps.T.x, k.ps.T.y, k.ps.T.z, __uint_as_float(k.ps.rng.s));
  c[3] = make_float4(k.ps.color.x, k.ps.color.y, k.ps.color.z, __uint_as_float((uint32_t)k.what | (k.ps.bounced ? 4u : 0u) | (k.alive ? 8u : 0u)));
  c[4] = make_float4(k.acc.x, k.acc.y, k.acc.z, __uint_as_float(k.pixp));
  c[5] = make_float4(__uint_as_float(k.s), __uint_as_float(k.s_end), __uint_as_float(k.slot), k.sh_len);
  if (k.what == WS_SHADOW) {
    c[6] = make_float4(k.ext_o.x, k.ext_o.y, k.ext_o.z, __int_as_float(k.sh_light));
    c[7] = make_float4(k.ext_d.x, k.ext_d.y, k.ext_d.z, 0.0f);
    c[8] = make_float4(k.contrib.x, k.contrib.y, k.contrib.z, 0.0f);
  }
  if (trav) c[9] = make_float4(__uint_as_float(root_lf), __uint_as_float(root_cnt), __int_as_float(-1), __uint_as_float(0u));
}
WPT_DEV void ctx_load(const float4* c, WCur& k) {
  float4 a = c[0], b = c[1], t = c[2], col = c[3], ac = c[4], m = c[5];
  k.o = xyz(a); k.res_t = a.w; k.d = xyz(b); k.res_id = __float_as_int(b.w);
  k.ps.T = xyz(t); k.ps.rng.s = __float_as_uint(t.w);
  k.ps.color = xyz(col);
  uint32_t f = __float_as_uint(col.w);
  k.what = (int)(f & 3u); k.ps.bounced = (f & 4u) != 0; k.alive = (f & 8u) != 0;
  k.acc = xyz(ac); k.pixp = __float_as_uint(ac.w);
  k.s = __float_as_uint(m.x); k.s_end = __float_as_uint(m.y); k.slot = __float_as_uint(m.z); k.sh_len = m.w;
  if (k.what == WS_SHADOW) {
    float4 e = c[6], g = c[7], h = c[8];
    k.ext_o = xyz(e); k.sh_light = __float_as_int(e.w); k.ext_d = xyz(g); k.contrib = xyz(h);
  }
}

// Scene::trace_g up to the point where the BVH has to be traversed (scene.rs:162-184): the infinite shapes, then the
// root guard (BVH2, scene.rs:191-212) or the root node's four child boxes (BVH4 has no root box test, scene.rs:292-342;
// a ray none of whose children survives ends here with the root's one visit). Returns true if the ray has to be queued:
// then (res_t, res_id) = (bound, plane-hit id), the job's input; else they are the final result of the ray.
template <int BVH, int KIND>
WPT_DEV bool wp_begin(const DScene& sc, F3 o, F3 d, float* res_t, int* res_id) {
  Ray ray = make_ray(o, d);
  Trav tv;
  bool enter = trav_begin<BVH, KIND>(sc, ray, tv);
  if (BVH == 4) {
    const float4* p = reinterpret_cast<const float4*>(sc.nodes4);
    float4 x0 = __ldg(p), y0 = __ldg(p + 1), z0 = __ldg(p + 2), x1 = __ldg(p + 3), y1 = __ldg(p + 4), z1 = __ldg(p + 5);
    uint32_t nc = __ldg(reinterpret_cast<const uint32_t*>(p + 7));
    float d0 = box_hit_x4(x0.x, y0.x, z0.x, x1.x, y1.x, z1.x, ray), d1 = box_hit_x4(x0.y, y0.y, z0.y, x1.y, y1.y, z1.y, ray);
    float d2 = box_hit_x4(x0.z, y0.z, z0.z, x1.z, y1.z, z1.z, ray), d3 = box_hit_x4(x0.w, y0.w, z0.w, x1.w, y1.w, z1.w, ray);
    enter = (nc > 0 && d0 >= 0.0f && !(d0 > tv.bound)) || (nc > 1 && d1 >= 0.0f && !(d1 > tv.bound)) ||
            (nc > 2 && d2 >= 0.0f && !(d2 > tv.bound)) || (nc > 3 && d3 >= 0.0f && !(d3 > tv.bound));
  }
  *res_t = enter ? tv.bound : tv.inf_t;
  *res_id = tv.inf_id;
  return enter;
}

template <int BVH, int KIND, int MINB, int RT>
__global__ void __launch_bounds__(WP_THREADS, MINB) k_wpool(MegaParams P) {
  const DScene& sc = P.rp.scene;
  const unsigned FULL = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  __shared__ uint8_t s_lq[WP_WARPS][WP_QCAP], s_tq[WP_WARPS][WP_QCAP];
  __shared__ unsigned int s_cnt[4];
  uint8_t* lq = s_lq[wib]; uint8_t* tq = s_tq[wib];
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const uint32_t C = P.pool_ctx;
  float4* pool = P.pool + (size_t)(blockIdx.x * WP_WARPS + wib) * C * WP_CTX_F4;
  const uint32_t nslots = P.nslots_dev ? *P.nslots_dev : P.nslots;
  uint32_t stack_n[WPT_STACK]; float stack_d[WPT_STACK];
  // warp-uniform state
  uint32_t lq_head = 0, lq_n = 0, tq_head = 0, tq_n = 0, created = 0;
  uint32_t chunk_next = 0, chunk_end = 0; bool fresh_left = true;
  uint32_t n_rays = 0, n_guard = 0, n_paths = 0;
  int jid = -1;   // the context this lane traverses for (kept across logic runs while the job is suspended)
#ifdef WP_INSTR
  unsigned long long i_runs = 0, i_pass = 0, i_have = 0, i_shade = 0, i_bursts = 0, i_iter = 0, i_run = 0, i_susp = 0, i_store = 0, i_load = 0, i_leave_hi = 0, i_leave_dry = 0;
#endif

  for (;;) {
    // ================================================================== logic run
    {
      WCur k; bool have = false; uint32_t cid = 0;
      k.what = WS_GEN; k.s = k.s_end = 0; k.alive = false; k.ps.bounced = false;
#ifdef WP_INSTR
      i_runs++;
#endif
      for (;;) {
        // ---- refill 1: contexts whose ray is done
        const unsigned need = __ballot_sync(FULL, !have);
        if (need && lq_n) {
          const uint32_t take = min((uint32_t)__popc(need), lq_n), r = (uint32_t)__popc(need & lt);
          if (!have && r < take) { cid = lq[(lq_head + r) & (WP_QCAP - 1u)]; ctx_load(pool + cid * WP_CTX_F4, k); have = true; }
#ifdef WP_INSTR
          i_load += take;
#endif
          lq_head = (lq_head + take) & (WP_QCAP - 1u); lq_n -= take;
        }
        // ---- refill 2: fresh segments for contexts whose segment is done and for new contexts
        const bool seg_done = have && k.what == WS_GEN && k.s >= k.s_end;
        if (seg_done) {   // RenderTarget::write of the segment (contract B10)
          const uint32_t px = k.pixp & 0xFFFFu, py = k.pixp >> 16, pix = py * P.rp.W + px;
          if (P.nseg > 1 || P.seg_list) P.seg_buf[k.slot] = make_float4(k.acc.x, k.acc.y, k.acc.z, 0.0f);
          else { float4 a = P.accum[pix]; P.accum[pix] = make_float4(a.x + k.acc.x, a.y + k.acc.y, a.z + k.acc.z, __uint_as_float(k.s)); }
        }
        const unsigned m_new = __ballot_sync(FULL, !have);
        const uint32_t allow_new = fresh_left ? min((uint32_t)__popc(m_new), C - created) : 0u;
        const bool cand = fresh_left && (seg_done || (!have && (uint32_t)__popc(m_new & lt) < allow_new));
        const unsigned m_cand = __ballot_sync(FULL, cand);
        if (m_cand) {
          const uint32_t want = (uint32_t)__popc(m_cand), avail = chunk_end - chunk_next;
          uint32_t base2 = 0, end2 = 0;
          if (want > avail) {
            uint32_t b = 0;
            if (lane == 0) b = atomicAdd(P.work_counter, P.chunk);
            b = __shfl_sync(FULL, b, 0);
            if (b >= nslots) fresh_left = false; else { base2 = b; end2 = min(b + P.chunk, nslots); }
          }
          bool got = false;
          if (cand) {
            const uint32_t r = (uint32_t)__popc(m_cand & lt);
            const uint32_t idx = r < avail ? chunk_next + r : base2 + (r - avail);
            got = r < avail || idx < end2;
            if (got) {
              uint32_t pslot = idx, j = 0;
              if (P.seg_list) { uint32_t e = P.seg_list[idx]; pslot = e >> 3; j = e & 7u; }
              else if (P.nseg > 1) { pslot = idx / P.nseg; j = idx - pslot * P.nseg; }
              const uint32_t pix = P.pixel[pslot];
              const uint32_t spp = P.spp_per_slot ? P.spp_per_slot[pslot] : P.uniform_spp;
              const uint32_t s0 = __float_as_uint(P.accum[pix].w);   // samples accumulated so far = next sample index
              const uint32_t b = min(j * P.seg_len, spp);
              const uint32_t py = pix / P.rp.W, px = pix - py * P.rp.W;
              k.slot = idx; k.pixp = px | (py << 16);
              k.s = s0 + b; k.s_end = s0 + min(b + P.seg_len, spp);
              k.acc = f3(0.0f, 0.0f, 0.0f); k.what = WS_GEN;
              if (!have) cid = created + (uint32_t)__popc(m_new & lt);
            }
          }
          if (want <= avail) chunk_next += want;
          else { const uint32_t used2 = min(want - avail, end2 - base2); chunk_next = base2 + used2; chunk_end = end2; }
          created += (uint32_t)__popc(__ballot_sync(FULL, cand && got && !have));
          have = cand ? got : have;   // a context that finds no more work dies
        } else if (seg_done) have = false;
        const unsigned m_have = __ballot_sync(FULL, have);
        if (!m_have) break;
#ifdef WP_INSTR
        i_pass++; i_have += __popc(m_have);
#endif
        // ---- half A: shadow-ray results and camera rays
        bool start = false, park = false, finish_a = false;
        if (have && k.what == WS_SHADOW) {   // Scene::shadow_ray, scene.rs:114-132
          const bool occluded = k.res_id >= 0 && k.res_t < k.sh_len && k.res_id != k.sh_light;
          if (!occluded) k.ps.color = k.ps.color + k.contrib;
          if (k.alive) { k.o = k.ext_o; k.d = k.ext_d; k.what = WS_EXTEND; start = true; }
          else { k.acc = k.acc + k.ps.color; k.s += 1; k.what = WS_GEN; finish_a = true; }   // RenderTarget::write, render_target.rs:55-58
        }
        if (have && !start && k.what == WS_GEN && k.s < k.s_end) {   // tracer.rs:176-196 — sample s of this pixel on its own stream
          const uint32_t px = k.pixp & 0xFFFFu, py = k.pixp >> 16;
          k.ps.rng.s = stream_seed(py * P.rp.W + px, k.s, STREAM_PATH, P.rp.base_seed);
          const float j1 = k.ps.rng.f32();
          const float j2 = k.ps.rng.f32();
          const Ray cr = camera_ray(P.rp.cam, px, py, j1, j2);
          k.o = cr.o; k.d = cr.d;
          k.ps.color = f3(0, 0, 0); k.ps.T = f3(1.0f, 1.0f, 1.0f); k.ps.bounced = false;
          k.what = WS_EXTEND; start = true;
        }
        if (start) park = wp_begin<BVH, KIND>(sc, k.o, k.d, &k.res_t, &k.res_id);
        {
          const unsigned ms = __ballot_sync(FULL, start), mp = __ballot_sync(FULL, start && park);
          n_rays += (uint32_t)__popc(ms); n_guard += (uint32_t)__popc(BVH == 4 ? (ms & ~mp) : ms);
          n_paths += (uint32_t)__popc(__ballot_sync(FULL, finish_a));
        }
        // ---- half B: shade the hit of a camera / bounce ray
        start = false;
        bool finish = false;
#ifdef WP_INSTR
        i_shade += __popc(__ballot_sync(FULL, have && !park && k.what == WS_EXTEND));
#endif
        if (have && !park && k.what == WS_EXTEND) {
          Ray ray; ray.o = k.o; ray.d = k.d;
          ray.inv = KIND == K_SIMPLE ? f3(0.0f, 0.0f, 0.0f) : f3(1.0f / k.d.x, 1.0f / k.d.y, 1.0f / k.d.z);
          ShadeOut so;
          shade_hit<KIND, RT>(P.rp, ray, k.res_id, k.res_t, k.ps, so);
          if (so.finished) finish = true;
          else if (so.shadow) {
            k.ext_o = so.next_o; k.ext_d = so.next_d; k.contrib = so.contrib; k.sh_len = so.sh_len; k.sh_light = so.sh_light;
            k.alive = so.survive;
            k.o = so.sh_o; k.d = so.sh_d; k.what = WS_SHADOW; start = true;
          } else if (so.survive) { k.o = so.next_o; k.d = so.next_d; start = true; }
          else finish = true;
          if (finish) { k.acc = k.acc + k.ps.color; k.s += 1; k.what = WS_GEN; }   // RenderTarget::write, render_target.rs:55-58
        }
        bool park_b = false;
        if (start) park_b = wp_begin<BVH, KIND>(sc, k.o, k.d, &k.res_t, &k.res_id);
        {
          const unsigned ms = __ballot_sync(FULL, start), mp = __ballot_sync(FULL, start && park_b);
          n_rays += (uint32_t)__popc(ms); n_guard += (uint32_t)__popc(BVH == 4 ? (ms & ~mp) : ms);
          n_paths += (uint32_t)__popc(__ballot_sync(FULL, finish));
        }
        park = park || park_b;
        // ---- leave the run when a warp's worth of rays waits (or the logic side runs dry); park
        const unsigned m_park = __ballot_sync(FULL, park);
        const uint32_t n_wait = tq_n + (uint32_t)__popc(m_park) + (uint32_t)__popc(__ballot_sync(FULL, jid >= 0));
        const uint32_t n_left = (uint32_t)__popc(m_have & ~m_park);
        const bool can_refill = lq_n > 0 || (fresh_left && created < C);
        const bool leave = n_wait >= P.t_hi || (n_wait > 0 && n_left < P.t_switch && !can_refill);
        const bool st_l = leave && have && !park;
#ifdef WP_INSTR
        i_store += __popc(__ballot_sync(FULL, park || st_l)); if (leave) { if (n_wait >= P.t_hi) i_leave_hi++; else i_leave_dry++; }
#endif
        if (park || st_l) {
          uint32_t rlf = 0, rcnt = 0;
          if (BVH == 2) { float4 rb = __ldg(reinterpret_cast<const float4*>(sc.nodes2) + 1); rlf = __float_as_uint(rb.z); rcnt = __float_as_uint(rb.w); }
          ctx_store(pool + cid * WP_CTX_F4, k, park, rlf, rcnt);
          have = false;
        }
        if (m_park) {
          if (park) tq[(tq_head + tq_n + (uint32_t)__popc(m_park & lt)) & (WP_QCAP - 1u)] = (uint8_t)cid;
          tq_n += (uint32_t)__popc(m_park);
        }
        if (leave) {
          const unsigned m_l = __ballot_sync(FULL, st_l);
          if (st_l) lq[(lq_head + lq_n + (uint32_t)__popc(m_l & lt)) & (WP_QCAP - 1u)] = (uint8_t)cid;
          lq_n += (uint32_t)__popc(m_l);
          break;
        }
      }
    }
    __syncwarp();
    const unsigned m_susp = __ballot_sync(FULL, jid >= 0);
    if (!tq_n && !m_susp) { if (!lq_n) break; else continue; }   // nothing to traverse: done, or back to logic
    // ================================================================== traversal burst
    {
      Ray ray = make_ray(f3(0, 0, 0), f3(1, 1, 1));
      Trav tv; tv.lf = tv.cnt = 0; tv.sp = 0; tv.bound = 0; tv.best_id = -1; tv.inf_t = 0; tv.inf_id = -1; tv.visits = tv.prims = 0;
      bool fin = false, need_load = jid >= 0;   // a suspended job is resumed by the lane that holds its stack
      int n_idle = 32 - __popc(m_susp);          // lanes without a job (warp-uniform, kept up to date at the refill site)
#ifdef WP_INSTR
      i_bursts++;
#endif
      for (;;) {
        unsigned run = __ballot_sync(FULL, jid >= 0 && !fin && !need_load);
#ifdef WP_INSTR
        i_iter++; i_run += __popc(run);
#endif
        int n_notrun = 32 - __popc(run);   // idle lanes + lanes whose ray is done but not flushed yet
        if (!run || (n_notrun >= (int)P.t_refill && (n_notrun > n_idle || tq_n > 0))) {
          // ---- flush finished rays, refill idle lanes
          const unsigned m_fin = __ballot_sync(FULL, fin);
          if (fin) {
            float4* c = pool + (uint32_t)jid * WP_CTX_F4;
            if (tv.best_id >= 0) { c[0].w = tv.bound; c[1].w = __int_as_float(tv.best_id); }   // closest (scene.rs:406-422): a BVH hit wins
            atomicAdd(&s_cnt[1], tv.visits); if (tv.prims) atomicAdd(&s_cnt[3], tv.prims);
            lq[(lq_head + lq_n + (uint32_t)__popc(m_fin & lt)) & (WP_QCAP - 1u)] = (uint8_t)jid;
            jid = -1; fin = false;
          }
          lq_n += (uint32_t)__popc(m_fin);
          const unsigned idle = __ballot_sync(FULL, jid < 0);
          if (idle && tq_n) {
            const uint32_t take = min((uint32_t)__popc(idle), tq_n), r = (uint32_t)__popc(idle & lt);
            if (jid < 0 && r < take) { jid = (int)tq[(tq_head + r) & (WP_QCAP - 1u)]; need_load = true; }
            tq_head = (tq_head + take) & (WP_QCAP - 1u); tq_n -= take;
          }
          if (need_load) {
            const float4* c = pool + (uint32_t)jid * WP_CTX_F4;
            const float4 a = c[0], b = c[1], t9 = c[9];
            ray = make_ray(xyz(a), xyz(b));
            tv.bound = a.w; tv.lf = __float_as_uint(t9.x); tv.cnt = __float_as_uint(t9.y); tv.best_id = __float_as_int(t9.z); tv.sp = (int)__float_as_uint(t9.w);
            tv.visits = 0; tv.prims = 0;
            need_load = false;
          }
          run = __ballot_sync(FULL, jid >= 0);
          n_idle = n_notrun = 32 - __popc(run);
          if (!run) break;
        }
        if (!tq_n && __popc(run) < (int)P.t_lo && (lq_n > 0 || n_notrun > n_idle || (fresh_left && created < C))) {
          // ---- few rays left and the logic side has work: flush, suspend the rest, leave
          const unsigned m_fin = __ballot_sync(FULL, fin);
#ifdef WP_INSTR
          i_susp += __popc(run & ~m_fin);
#endif
          if (jid >= 0) {
            float4* c = pool + (uint32_t)jid * WP_CTX_F4;
            atomicAdd(&s_cnt[1], tv.visits); if (tv.prims) atomicAdd(&s_cnt[3], tv.prims);
            if (fin) {
              if (tv.best_id >= 0) { c[0].w = tv.bound; c[1].w = __int_as_float(tv.best_id); }
              lq[(lq_head + lq_n + (uint32_t)__popc(m_fin & lt)) & (WP_QCAP - 1u)] = (uint8_t)jid;
              jid = -1;
            } else {
              c[0].w = tv.bound;
              c[9] = make_float4(__uint_as_float(tv.lf), __uint_as_float(tv.cnt), __int_as_float(tv.best_id), __uint_as_float((uint32_t)tv.sp));
            }
          }
          lq_n += (uint32_t)__popc(m_fin);
          break;
        }
        // ---- one voted step (k_mega's burst body): inner-node steps while more than 1 / t_inner of the running lanes are
        // at inner nodes, else a leaf step with every lane that waits at a leaf, then the pop
        const bool tr = jid >= 0 && !fin;
        const bool leaf = tr && trav_at_leaf<BVH>(tv);
        const int n_inner = __popc(__ballot_sync(FULL, tr && !leaf)), n_run = __popc(run);
        if (n_inner == n_run || n_inner * (int)P.t_inner > n_run) {
#pragma unroll 1
          for (uint32_t rep = 0; rep < P.inner_reps; rep++) {
            bool np = false;
            if (jid >= 0 && !fin && !trav_at_leaf<BVH>(tv)) np = trav_inner<BVH>(sc, ray, tv, stack_n, stack_d);
            if (np && !trav_pop<BVH>(sc, tv, stack_n, stack_d)) fin = true;
          }
        } else if (leaf) {
          trav_leaf<BVH, KIND>(sc, ray, tv);
          if (!trav_pop<BVH>(sc, tv, stack_n, stack_d)) fin = true;
        }
      }
    }
    __syncwarp();
  }
  // ---- counters: rays, node visits (one per root guard + the jobs'), paths, leaf primitive tests
  if (lane == 0) { atomicAdd(&s_cnt[0], n_rays); atomicAdd(&s_cnt[1], n_guard); atomicAdd(&s_cnt[2], n_paths); }
#ifdef WP_INSTR
  if (lane == 0) { unsigned long long v[12] = {i_runs, i_pass, i_have, i_shade, i_bursts, i_iter, i_run, i_susp, i_store, i_load, i_leave_hi, i_leave_dry}; for (int i = 0; i < 12; i++) atomicAdd(&P.counters[4 + i], v[i]); }
#endif
  __syncthreads();
  if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(&P.counters[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

template <int BVH, int KIND, int RT>
static void launch_wpool_t(const MegaParams& P, int grid, int minb, cudaStream_t s) {
  if (minb >= 12) k_wpool<BVH, KIND, 12, RT><<<grid, WP_THREADS, 0, s>>>(P);
  else if (minb >= 8) k_wpool<BVH, KIND, 8, RT><<<grid, WP_THREADS, 0, s>>>(P);
  else k_wpool<BVH, KIND, 6, RT><<<grid, WP_THREADS, 0, s>>>(P);
}
template <int BVH, int KIND>
static void launch_wpool_rt(const MegaParams& P, int grid, int minb, cudaStream_t s) {
  switch (P.rp.render_type) {
    case 0: launch_wpool_t<BVH, KIND, 0>(P, grid, minb, s); break;
    case 2: launch_wpool_t<BVH, KIND, 2>(P, grid, minb, s); break;
    default: launch_wpool_t<BVH, KIND, 1>(P, grid, minb, s); break;
  }
}
uint32_t wpool_warps(int grid) { return (uint32_t)grid * WP_WARPS; }
size_t wpool_ctx_bytes() { return WP_CTX_F4 * sizeof(float4); }
void launch_wpool(const MegaParams& P, int grid, int minb, cudaStream_t s) {
  if (!P.nslots) return;
  const bool b4 = P.rp.scene.bvh_kind == 4;
  if (P.scene_kind == K_SIMPLE) { if (b4) launch_wpool_rt<4, K_SIMPLE>(P, grid, minb, s); else launch_wpool_rt<2, K_SIMPLE>(P, grid, minb, s); }
  else { if (b4) launch_wpool_rt<4, K_EXT>(P, grid, minb, s); else launch_wpool_rt<2, K_EXT>(P, grid, minb, s); }
}

}  // namespace wpt
