// ORACLE — TEST INFRASTRUCTURE ONLY. PARITY UNPINNED. See ref_core.h.
// C entry points (ctypes) over the CPU restatement. Loaded only by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
#include "ref_session.h"
#include <cstdio>

using namespace ref;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH(ret) } catch (const std::exception& e) { g_err = e.what(); return ret; }

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_create(uint32_t w, uint32_t h, uint32_t scene_id, float cx, float cy, float cz, float rx, float ry, int bvh4) {
  ORC_TRY return new Session(w, h, scene_id, cx, cy, cz, rx, ry, bvh4 != 0); ORC_CATCH(nullptr)
}
void orc_destroy(void* s) { delete (Session*)s; }

// ---- wasm_interface.rs mirror (mode A)
int orc_update_scene(void* s, uint32_t id) { ORC_TRY ((Session*)s)->update_scene(id); return 0; ORC_CATCH(-1) }
int orc_update_settings(void* s, uint32_t lt, uint32_t rt, uint32_t la, uint32_t ra, uint32_t dbg) { ORC_TRY ((Session*)s)->update_settings(lt, rt, la, ra, dbg); return 0; ORC_CATCH(-1) }
int orc_update_viewport(void* s, uint32_t w, uint32_t h) { ORC_TRY ((Session*)s)->update_viewport(w, h); return 0; ORC_CATCH(-1) }
int orc_update_camera(void* s, float x, float y, float z, float rx, float ry) { ORC_TRY ((Session*)s)->update_camera(x, y, z, rx, ry); return 0; ORC_CATCH(-1) }
// allocate_texture + the client's copy in one call (wasm_interface.rs:335-352, worker.ts:182-190): packed RGB8, row-major
int orc_store_texture(void* s, uint32_t id, uint32_t w, uint32_t h, const uint8_t* rgb) {
  ORC_TRY Session* S = (Session*)s; S->textures[id].assign(rgb, rgb + (size_t)w * h * 3); S->tex_size[id] = std::make_pair(w, h); return 0; ORC_CATCH(-1)
}
int orc_allocate_mesh(void* s, uint32_t id, uint32_t nv) { ORC_TRY ((Session*)s)->allocate_mesh(id, nv); return 0; ORC_CATCH(-1) }
float* orc_mesh_vertices(void* s, uint32_t id) { ORC_TRY return ((Session*)s)->mesh_vertices(id); ORC_CATCH(nullptr) }
int orc_notify_mesh_loaded(void* s, uint32_t id) { ORC_TRY return ((Session*)s)->notify_mesh_loaded(id) ? 1 : 0; ORC_CATCH(-1) }
int orc_compute(void* s, uint64_t n) { ORC_TRY ((Session*)s)->compute((size_t)n); return 0; ORC_CATCH(-1) }
int orc_reset(void* s) { ORC_TRY ((Session*)s)->reset(); return 0; ORC_CATCH(-1) }
const uint8_t* orc_results(void* s, uint32_t show_sampling) {
  Session* S = (Session*)s;
  return show_sampling == 1 ? S->sampling_target->result.data() : S->target->result.data();
}
int orc_rebuild_bvh(void* s, int bvh4) { ORC_TRY ((Session*)s)->rebuild_bvh(bvh4 != 0); return 0; ORC_CATCH(-1) }
void orc_set_trig_a(void* s, int trig) { Session* S = (Session*)s; S->trig_a = (TrigMode)trig; S->left->trig = S->right->trig = (TrigMode)trig; }

// ---- mode B
int orc_mb_config(void* s, int type, int light_debug, int trig, uint32_t seed, uint64_t photon_target, uint32_t rx, uint32_t ry, uint32_t rw, uint32_t rh) {
  Session* S = (Session*)s;
  ORC_TRY
  if (type < 0 || type > 2) throw std::runtime_error("Invalid RenderType magic number");
  bool rebuild_photons = S->mb.base_seed != seed || S->mb.photon_target != photon_target;
  S->mb.type = (RenderType)type; S->mb.light_debug = light_debug != 0; S->mb.trig = (TrigMode)trig;
  S->mb.base_seed = seed; S->mb.photon_target = (size_t)photon_target;
  S->mb.rx = rx; S->mb.ry = ry; S->mb.rw = rw ? rw : S->W; S->mb.rh = rh ? rh : S->H;
  if (S->mb_round_left.size() != S->mb.rw * S->mb.rh) { S->mb_round_left.clear(); S->mb_adaptive_started = false; }
  if (rebuild_photons) { S->mb_photons.reset(); S->mb_photon_list.clear(); S->mb_shots = 0; }
  return 0;
  ORC_CATCH(-1)
}
int orc_mb_build_photons(void* s, uint32_t threads) { ORC_TRY ((Session*)s)->mb_build_photons(threads ? threads : 1); return 0; ORC_CATCH(-1) }
int orc_mb_render_exact(void* s, uint32_t spp, uint32_t threads) { ORC_TRY ((Session*)s)->mb_render_exact(spp, threads ? threads : 1); return 0; ORC_CATCH(-1) }
int64_t orc_mb_render_adaptive(void* s, uint64_t budget, uint32_t threads) { ORC_TRY return (int64_t)((Session*)s)->mb_render_adaptive(budget, threads ? threads : 1); ORC_CATCH(-1) }
int orc_mb_render_random(void* s, uint64_t ticks, uint32_t threads) { ORC_TRY ((Session*)s)->mb_render_random(ticks, threads ? threads : 1); return 0; ORC_CATCH(-1) }
int orc_mb_primary_probe(void* s, int32_t* ids, uint32_t* visits, float* dist) { ORC_TRY ((Session*)s)->mb_primary_probe(ids, visits, dist); return 0; ORC_CATCH(-1) }
int orc_mb_round_spp(void* s, uint32_t* out, uint64_t cap) {
  Session* S = (Session*)s;
  size_t n = std::min((size_t)cap, S->mb_round_spp.size());
  for (size_t i = 0; i < n; i++) out[i] = S->mb_round_spp[i];
  return (int)S->mb_round_spp.size();
}
// error map of the current region (mode-B reduction): mse[rw*rh], stats[3] = min, avg, max
int orc_mb_error_map(void* s, float* mse_out, float* stats) {
  Session* S = (Session*)s;
  ORC_TRY
  std::vector<float> mse; float mn, avg, mx;
  S->mb_error_stats(mse, &mn, &avg, &mx);
  if (mse_out) std::copy(mse.begin(), mse.end(), mse_out);
  stats[0] = mn; stats[1] = avg; stats[2] = mx;
  return 0;
  ORC_CATCH(-1)
}

// ---- read-backs
void orc_accum(void* s, float* rgb, uint32_t* counts) {
  Session* S = (Session*)s;
  size_t n = S->W * S->H;
  for (size_t i = 0; i < n; i++) {
    if (rgb) { rgb[i * 3] = S->target->acc_buffer[i].x; rgb[i * 3 + 1] = S->target->acc_buffer[i].y; rgb[i * 3 + 2] = S->target->acc_buffer[i].z; }
    if (counts) counts[i] = (uint32_t)S->target->acc_count[i];
  }
}
// out[0..7] = rays, paths, node_visits, photons_shot, photons_stored, prim_tests, 0, 0
void orc_stats(void* s, int which, uint64_t* out) {
  Session* S = (Session*)s;
  Stats st;
  if (which == 0) st = S->mb_stats;
  else if (which == 1) st = S->left->stats;
  else st = S->right->stats;
  out[0] = st.rays; out[1] = st.paths; out[2] = st.node_visits; out[3] = st.photons_shot; out[4] = st.photons_stored; out[5] = st.prim_tests; out[6] = out[7] = 0;
}
// scene introspection: info[0]=num_shapes, [1]=num_inf, [2]=num_lights, [3]=bvh2 nodes, [4]=bvh2 depth, [5]=bvh4 nodes(array len), [6]=bvh kind, [7]=bvh4 depth
void orc_scene_info(void* s, uint64_t* info) {
  Session* S = (Session*)s;
  const Scene& sc = *S->scene;
  info[0] = sc.shapes.size(); info[1] = sc.num_inf; info[2] = sc.lights.size(); info[3] = sc.bvh2.size();
  info[4] = sc.shapes.size() > sc.num_inf ? bvh_depth(sc.bvh2) : 0; info[5] = sc.bvh4.size(); info[6] = sc.bvh_kind;
  info[7] = sc.bvh_kind == 4 ? bvh4_depth(sc.bvh4) : 0;
}
// BVH2 node array: bounds[n*6] (xmin,ymin,zmin,xmax,ymax,zmax), lf[n], cnt[n]
void orc_bvh2(void* s, float* bounds, uint32_t* lf, uint32_t* cnt) {
  const Scene& sc = *((Session*)s)->scene;
  for (size_t i = 0; i < sc.bvh2.size(); i++) {
    const AABB& b = sc.bvh2[i].bounds;
    float v[6] = {b.x_min, b.y_min, b.z_min, b.x_max, b.y_max, b.z_max};
    std::copy(v, v + 6, bounds + i * 6);
    lf[i] = sc.bvh2[i].left_first; cnt[i] = sc.bvh2[i].count;
  }
}
// BVH4 node array: bounds[n*24] (child-major: 4 x 6 floats), children[n*4], num_children[n]
void orc_bvh4(void* s, float* bounds, int32_t* children, uint32_t* nc) {
  const Scene& sc = *((Session*)s)->scene;
  for (size_t i = 0; i < sc.bvh4.size(); i++) {
    for (int c = 0; c < 4; c++) {
      const AABB& b = sc.bvh4[i].child_bounds[c];
      float v[6] = {b.x_min, b.y_min, b.z_min, b.x_max, b.y_max, b.z_max};
      std::copy(v, v + 6, bounds + i * 24 + c * 6);
      children[i * 4 + c] = sc.bvh4[i].children[c];
    }
    nc[i] = sc.bvh4[i].num_children;
  }
}
// per shape (final order): source index, type; lights[]: shape index per light
void orc_shape_order(void* s, int32_t* source_index, int32_t* type) {
  const Scene& sc = *((Session*)s)->scene;
  for (size_t i = 0; i < sc.shapes.size(); i++) { source_index[i] = sc.shapes[i].source_index; type[i] = (int)sc.shapes[i].type; }
}
void orc_lights(void* s, uint32_t* out) {
  const Scene& sc = *((Session*)s)->scene;
  for (size_t i = 0; i < sc.lights.size(); i++) out[i] = (uint32_t)sc.lights[i];
}
int orc_verify_bvh(void* s) {
  const Scene& sc = *((Session*)s)->scene;
  bool ok = verify_bvh(sc.shapes, sc.num_inf, sc.bvh2);
  if (sc.bvh_kind == 4) ok = ok && verify_bvh4(sc.shapes, sc.num_inf, sc.bvh4);
  return ok ? 1 : 0;
}

// generic ray batch through trace_g (+ full hit normal): o[n*3], d[n*3] -> ids, dist, visits, normal[n*3]
int orc_trace_rays(void* s, const float* o, const float* d, uint64_t n, int32_t* ids, float* dist, uint32_t* visits, float* normals) {
  Session* S = (Session*)s;
  ORC_TRY
  for (uint64_t i = 0; i < n; i++) {
    Ray ray(Vec3(o[i * 3], o[i * 3 + 1], o[i * 3 + 2]), Vec3(d[i * 3], d[i * 3 + 1], d[i * 3 + 2]));
    GHit g; size_t v = S->scene->trace_g(ray, &g);
    ids[i] = g.some ? (int32_t)g.shape : -1;
    dist[i] = g.some ? g.dis : INF_F;
    visits[i] = (uint32_t)v;
    if (normals) {
      Hit h; bool ok = g.some && S->scene->shapes[g.shape].trace(ray, &h);
      normals[i * 3] = ok ? h.normal.x : 0.0f; normals[i * 3 + 1] = ok ? h.normal.y : 0.0f; normals[i * 3 + 2] = ok ? h.normal.z : 0.0f;
    }
  }
  return 0;
  ORC_CATCH(-1)
}

// ---- photons (mode B)
uint64_t orc_mb_photon_count(void* s) { return ((Session*)s)->mb_photon_list.size(); }
uint64_t orc_mb_photon_shots(void* s) { return ((Session*)s)->mb_shots; }
void orc_mb_photon_list(void* s, uint32_t* light, float* loc, float* w) {
  Session* S = (Session*)s;
  for (size_t i = 0; i < S->mb_photon_list.size(); i++) {
    const PhotonRec& p = S->mb_photon_list[i];
    light[i] = (uint32_t)p.light; loc[i * 3] = p.loc.x; loc[i * 3 + 1] = p.loc.y; loc[i * 3 + 2] = p.loc.z; w[i] = p.w;
  }
}
static void flatten_tree(Octree& o, uint32_t depth, std::vector<uint32_t>& meta, std::vector<float>& cum, std::vector<float>& bins) {
  o.cdf.recheck_cdf();
  meta.push_back(depth); meta.push_back(o.is_node ? 1u : 0u); meta.push_back((uint32_t)o.values.size());
  cum.insert(cum.end(), o.cdf.cum_bins.begin(), o.cdf.cum_bins.end());
  bins.insert(bins.end(), o.cdf.bins.begin(), o.cdf.bins.end());
  for (auto& c : o.children) flatten_tree(c, depth + 1, meta, cum, bins);
}
// photon tree in DFS pre-order (children in octant order): meta[n*3] = depth,is_node,leaf photons;
// cum[n*L], bins[n*L]. Call with null buffers to get n.
uint64_t orc_mb_photon_tree(void* s, uint32_t* meta, float* cum, float* bins) {
  Session* S = (Session*)s;
  if (!S->mb_photons) return 0;
  std::vector<uint32_t> m; std::vector<float> c, b;
  flatten_tree(S->mb_photons->root, 0, m, c, b);
  if (meta) std::copy(m.begin(), m.end(), meta);
  if (cum) std::copy(c.begin(), c.end(), cum);
  if (bins) std::copy(b.begin(), b.end(), bins);
  return m.size() / 3;
}
// PNEE light choice for a batch of points, each with its own stream seed: -> light, pdf
int orc_mb_photon_sample(void* s, const float* pts, const uint32_t* seeds, uint64_t n, uint32_t* light, float* pdf) {
  Session* S = (Session*)s;
  ORC_TRY
  if (!S->mb_photons) throw std::runtime_error("photon tree not built");
  for (uint64_t i = 0; i < n; i++) {
    Rng r(seeds[i]); size_t l; float p;
    S->mb_photons->sample(r, Vec3(pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]), &l, &p);
    light[i] = (uint32_t)l; pdf[i] = p;
  }
  return 0;
  ORC_CATCH(-1)
}

// ---- known-answer helpers
void orc_rng_u32(uint32_t seed, uint32_t n, uint32_t* out) { Rng r(seed); for (uint32_t i = 0; i < n; i++) out[i] = r.next_u32(); }
void orc_rng_f32(uint32_t seed, uint32_t n, float* out) { Rng r(seed); for (uint32_t i = 0; i < n; i++) out[i] = r.next(); }
uint32_t orc_rng_range(uint32_t seed, uint32_t n, uint32_t lo, uint32_t hi, uint32_t* out) { Rng r(seed); for (uint32_t i = 0; i < n; i++) out[i] = (uint32_t)r.next_in_range(lo, hi); return r.state; }
uint32_t orc_stream_seed(uint32_t index, uint32_t sample, uint32_t stream, uint32_t base) { return stream_seed(index, sample, stream, base); }
uint32_t orc_museum_colors(int32_t* order27) { std::vector<int> o; uint32_t st; museum_shapes(&o, &st); for (size_t i = 0; i < o.size() && i < 27; i++) order27[i] = o[i]; return st; }
void orc_shared_sincos(const float* a, uint32_t n, float* s, float* c) { for (uint32_t i = 0; i < n; i++) shared_sincos(a[i], &s[i], &c[i]); }
void orc_shared_exp_neg(const float* x, uint32_t n, float* out) { for (uint32_t i = 0; i < n; i++) out[i] = shared_exp_neg(x[i]); }
void orc_hemisphere(uint32_t seed, uint32_t n, const float* normal, float* out) {
  Rng r(seed); Vec3 nn(normal[0], normal[1], normal[2]);
  for (uint32_t i = 0; i < n; i++) { Vec3 v = r.next_hemisphere(nn); out[i * 3] = v.x; out[i * 3 + 1] = v.y; out[i * 3 + 2] = v.z; }
}
// EmpiricalPDF (main.rs:54-81 experiment): set bins, draw n samples, histogram
void orc_empirical_pdf(const float* bins, uint32_t nb, uint32_t seed, uint32_t n, uint32_t* hist, float* probs) {
  EmpiricalPDF pdf(nb, false);
  for (uint32_t i = 0; i < nb; i++) pdf.set(i, bins[i]);
  Rng r(seed);
  for (uint32_t i = 0; i < nb; i++) hist[i] = 0;
  for (uint32_t i = 0; i < n; i++) hist[pdf.sample(r)]++;
  for (uint32_t i = 0; i < nb; i++) probs[i] = pdf.bin_prob(i);
}
// quartic solver: coef[5] = a4..a0 -> roots (ascending), returns count
void orc_shared_f64(int which, const double* x, uint32_t n, double* out) { for (uint32_t i = 0; i < n; i++) out[i] = which == 0 ? shared_cos64(x[i]) : which == 1 ? shared_acos64(x[i]) : shared_cbrt64(x[i]); }
int orc_quartic(const double* coef, double* roots_out) { Roots r = roots_quartic(coef[0], coef[1], coef[2], coef[3], coef[4]); for (int i = 0; i < r.n; i++) roots_out[i] = r.v[i]; return r.n; }
// OBJ text -> expanded vertices (9 floats per triangle); returns float count, fills up to cap
int64_t orc_parse_obj(const char* text, uint64_t len, int client_scale, float* out, uint64_t cap) {
  ORC_TRY
  std::vector<float> v = parse_obj(std::string(text, len), client_scale != 0);
  for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
  return (int64_t)v.size();
  ORC_CATCH(-1)
}

}  // extern "C"
