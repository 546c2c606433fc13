// wpt_render — progressive driver over the C ABI of libwpt.so, written against include/wpt.h only.
//
// It replays what the reference's worker does with the WASM module (src_ts/worker/worker.ts):
//   handleInit        -> wpt_init                                   (worker.ts:98-140)
//   handleStoreMesh   -> allocate_mesh / mesh_vertices / notify     (worker.ts:171-179)
//   run()             -> update_camera? update_viewport? compute(numRaysPerTick) results()
//                        with numRaysPerTick rescaled so that one tick takes ~50 ms
//                        (worker.ts:55-95), until the time / tick / sample budget is spent
//   pause / resume    -> the loop simply stops between two ticks     (worker.ts:158-168)
//   update_view_type  -> results(1): the sampling-strategy view      (worker.ts:150-160)
// and stores the RGBA8 frame as a binary PPM or an (uncompressed-deflate) PNG.
//
//   wpt_render [--scene 0|2] [--obj file.obj] [--size WxH] [--left T] [--right T]
//              [--left-adaptive 0|1] [--right-adaptive 0|1] [--light-debug 0|1]
//              [--seconds S] [--ticks N] [--samples N] [--tick-ms 50] [--sampling-view]
//              [--out frame.png|frame.ppm] [--camera x,y,z,rx,ry] [--quiet]
//
// Frame mode — one full-frame render on 1..N GPUs through the library's native NCCL plane (no Python anywhere):
//   wpt_render --frame-spp S [--type 0|1|2] [--bvh 2|4] [--adaptive] [--photons N]
//              [--world N --rank R --nccl-id-file PATH] [--device D]        (defaults: WORLD_SIZE / RANK / LOCAL_RANK)
// One process per GPU, e.g.  for r in 0 1; do tools/wpt_render --frame-spp 16 --world 2 --rank $r --nccl-id-file /tmp/id ... & done; wait
// (or under torchrun / mpirun, which set the environment variables). Rank 0 creates the NCCL unique id and writes it to
// the file, the others wait for it; every rank renders its 4-row bands, wpt_ctx_gather_frame() all-gathers the
// accumulators, rank 0 stores the image. The FNV-1a hash of the RGBA8 frame is printed: it equals the 1-GPU run's.
//
// Build: g++ -O2 -std=c++17 tools/wpt_render.cpp -Iinclude -Lwasm_pathtracer_b200 -lwpt -Wl,-rpath,'$ORIGIN/../wasm_pathtracer_b200' -o tools/wpt_render
#include <chrono>
#include <ctime>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "wpt.h"

static void die(const char* what) {
  std::fprintf(stderr, "wpt_render: %s: %s\n", what, wpt_last_error());
  std::exit(1);
}

// ---- PNG with stored (uncompressed) deflate blocks: no zlib needed
static uint32_t crc_table[256];
static void crc_init() {
  for (uint32_t n = 0; n < 256; n++) {
    uint32_t c = n;
    for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    crc_table[n] = c;
  }
}
static uint32_t crc32(const uint8_t* p, size_t n, uint32_t c = 0xFFFFFFFFu) {
  for (size_t i = 0; i < n; i++) c = crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c;
}
static void be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
static void chunk(std::vector<uint8_t>& out, const char tag[4], const std::vector<uint8_t>& data) {
  be32(out, (uint32_t)data.size());
  std::vector<uint8_t> td(tag, tag + 4);
  td.insert(td.end(), data.begin(), data.end());
  out.insert(out.end(), td.begin(), td.end());
  be32(out, crc32(td.data(), td.size()) ^ 0xFFFFFFFFu);
}
static bool write_png(const std::string& path, const uint8_t* rgba, uint32_t w, uint32_t h) {
  crc_init();
  std::vector<uint8_t> raw;
  raw.reserve((size_t)h * (w * 4 + 1));
  for (uint32_t y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), rgba + (size_t)y * w * 4, rgba + (size_t)(y + 1) * w * 4); }
  std::vector<uint8_t> z = {0x78, 0x01};
  uint32_t a = 1, b = 0;
  for (size_t i = 0; i < raw.size(); i++) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
  for (size_t off = 0; off < raw.size(); off += 65535) {
    size_t n = std::min<size_t>(65535, raw.size() - off);
    z.push_back(off + n == raw.size() ? 1 : 0);
    z.push_back(n & 0xFF); z.push_back(n >> 8); z.push_back(~n & 0xFF); z.push_back((~n >> 8) & 0xFF);
    z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
  }
  be32(z, (b << 16) | a);
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  std::vector<uint8_t> ihdr;
  be32(ihdr, w); be32(ihdr, h);
  ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  chunk(out, "IHDR", ihdr); chunk(out, "IDAT", z); chunk(out, "IEND", {});
  std::ofstream f(path, std::ios::binary);
  f.write((const char*)out.data(), (std::streamsize)out.size());
  return (bool)f;
}
static bool write_ppm(const std::string& path, const uint8_t* rgba, uint32_t w, uint32_t h) {
  std::ofstream f(path, std::ios::binary);
  f << "P6\n" << w << " " << h << "\n255\n";
  for (size_t i = 0; i < (size_t)w * h; i++) f.write((const char*)rgba + i * 4, 3);
  return (bool)f;
}

static uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}
static uint32_t env_u32(const char* name, uint32_t dflt) { const char* v = std::getenv(name); return v ? (uint32_t)std::atoi(v) : dflt; }

// Frame mode: handle API + native NCCL plane (include/wpt.h, "Native multi-GPU plane").
static int frame_mode(uint32_t scene, uint32_t w, uint32_t h, const float cam[5], const std::string& obj, uint32_t type, uint32_t bvh, bool adaptive, uint32_t spp,
                      uint64_t photons, uint32_t rank, uint32_t world, int device, const std::string& id_file, const std::string& out, bool quiet) {
  wpt_ctx* c = wpt_ctx_create(device, w, h, scene, cam[0], cam[1], cam[2], cam[3], cam[4]);
  if (!c) die("ctx_create");
  if (!obj.empty() && wpt_ctx_load_obj(c, 1, obj.c_str(), 1) < 0) die("load_obj");
  wpt_config cfg;
  if (wpt_ctx_get_config(c, &cfg) != 0) die("get_config");
  cfg.bvh_kind = bvh; cfg.render_type = type; if (photons) cfg.photon_target = photons;
  if (wpt_ctx_set_config(c, &cfg) != 0) die("set_config");
  uint8_t id[128] = {0};
  if (world > 1) {
    if (id_file.empty()) { std::fprintf(stderr, "wpt_render: --nccl-id-file is needed when --world > 1\n"); return 2; }
    if (rank == 0) {
      if (wpt_nccl_unique_id(id) != 0) die("nccl_unique_id");
      std::string tmp = id_file + ".tmp";
      { std::ofstream f(tmp, std::ios::binary); f.write((const char*)id, 128); }
      std::rename(tmp.c_str(), id_file.c_str());
    } else {
      for (int tries = 0;; tries++) {
        std::ifstream f(id_file, std::ios::binary);
        if (f && f.read((char*)id, 128)) break;
        if (tries > 6000) { std::fprintf(stderr, "wpt_render: no NCCL id in %s\n", id_file.c_str()); return 1; }
        struct timespec ts = {0, 10 * 1000 * 1000}; nanosleep(&ts, nullptr);
      }
    }
  }
  if (wpt_ctx_attach_nccl(c, id, rank, world) != 0) die("attach_nccl");
  using clock = std::chrono::steady_clock;
  auto t0 = clock::now();
  if (type == WPT_PNEE && wpt_ctx_build_photons(c) != 0) die("build_photons");
  wpt_ctx_synchronize(c);
  double warm = std::chrono::duration<double>(clock::now() - t0).count();
  t0 = clock::now();
  if (adaptive) { if (wpt_ctx_render_adaptive(c, (uint64_t)w * h * spp) < 0) die("render_adaptive"); }   // exchanges between the rounds
  else { if (wpt_ctx_render_exact(c, spp) != 0) die("render_exact"); if (wpt_ctx_gather_frame(c) != 0) die("gather_frame"); }
  const uint8_t* frame = wpt_ctx_results(c, 0);
  if (!frame) die("results");
  double secs = std::chrono::duration<double>(clock::now() - t0).count();
  uint64_t st[8] = {0};
  wpt_ctx_stats(c, st);
  std::printf("{\"rank\": %u, \"world\": %u, \"seconds\": %.4f, \"photon_warmup_s\": %.4f, \"rays\": %llu, \"paths\": %llu, \"node_visits\": %llu, \"frame_fnv1a\": \"%016llx\"}\n",
              rank, world, secs, warm, (unsigned long long)st[0], (unsigned long long)st[1], (unsigned long long)st[2], (unsigned long long)fnv1a(frame, (size_t)w * h * 4));
  int rc = 0;
  if (rank == 0 && !out.empty()) {
    bool ok = out.size() > 4 && out.substr(out.size() - 4) == ".ppm" ? write_ppm(out, frame, w, h) : write_png(out, frame, w, h);
    if (!ok) { std::fprintf(stderr, "wpt_render: cannot write %s\n", out.c_str()); rc = 1; }
  }
  (void)quiet;
  wpt_ctx_detach_nccl(c);
  wpt_ctx_destroy(c);
  return rc;
}

int main(int argc, char** argv) {
  uint32_t scene = WPT_SCENE_BUNNY, w = 512, h = 512;
  uint32_t f_spp = 0, f_type = WPT_NORMAL_NEE, f_bvh = 2, f_rank = env_u32("RANK", 0), f_world = env_u32("WORLD_SIZE", 1); bool f_adaptive = false;
  int f_device = (int)env_u32("LOCAL_RANK", 0); uint64_t f_photons = 0; std::string id_file;
  uint32_t left = WPT_NORMAL_NEE, right = WPT_PNEE, left_ad = 0, right_ad = 1, light_debug = 0;   // the reference's defaults (wasm_interface.rs:90-101)
  double seconds = 2.0, tick_ms = 50.0;
  uint64_t max_ticks = 0, max_samples = 0;
  bool sampling_view = false, quiet = false, cam_set = false;
  float cam[5] = {-0.9f, 5.4f, 0.4f, 0.58f, 0.0f};   // bunny camera, index.ts:158
  std::string obj, out = "frame.png";
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
    if (a == "--scene") scene = (uint32_t)std::atoi(next());
    else if (a == "--obj") obj = next();
    else if (a == "--size") { if (std::sscanf(next(), "%ux%u", &w, &h) != 2) { std::fprintf(stderr, "--size WxH\n"); return 2; } }
    else if (a == "--left") left = (uint32_t)std::atoi(next());
    else if (a == "--right") right = (uint32_t)std::atoi(next());
    else if (a == "--left-adaptive") left_ad = (uint32_t)std::atoi(next());
    else if (a == "--right-adaptive") right_ad = (uint32_t)std::atoi(next());
    else if (a == "--light-debug") light_debug = (uint32_t)std::atoi(next());
    else if (a == "--seconds") seconds = std::atof(next());
    else if (a == "--ticks") max_ticks = std::strtoull(next(), nullptr, 10);
    else if (a == "--samples") max_samples = std::strtoull(next(), nullptr, 10);
    else if (a == "--tick-ms") tick_ms = std::atof(next());
    else if (a == "--sampling-view") sampling_view = true;
    else if (a == "--quiet") quiet = true;
    else if (a == "--out") out = next();
    else if (a == "--frame-spp") f_spp = (uint32_t)std::atoi(next());
    else if (a == "--type") f_type = (uint32_t)std::atoi(next());
    else if (a == "--bvh") f_bvh = (uint32_t)std::atoi(next());
    else if (a == "--adaptive") f_adaptive = true;
    else if (a == "--photons") f_photons = std::strtoull(next(), nullptr, 10);
    else if (a == "--rank") f_rank = (uint32_t)std::atoi(next());
    else if (a == "--world") f_world = (uint32_t)std::atoi(next());
    else if (a == "--device") f_device = std::atoi(next());
    else if (a == "--nccl-id-file") id_file = next();
    else if (a == "--camera") { if (std::sscanf(next(), "%f,%f,%f,%f,%f", cam, cam + 1, cam + 2, cam + 3, cam + 4) != 5) { std::fprintf(stderr, "--camera x,y,z,rx,ry\n"); return 2; } cam_set = true; }
    else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }
  if (!cam_set && scene == WPT_SCENE_MUSEUM) { const float m[5] = {0.0f, 16.34f, -23.76f, 0.54f, 0.0f}; std::memcpy(cam, m, sizeof m); }   // index.ts:156

  if (f_spp) return frame_mode(scene, w, h, cam, obj, f_type, f_bvh, f_adaptive, f_spp, f_photons, f_rank, f_world, f_device, id_file, out, quiet);

  // handleInit
  wpt_init(w, h, scene, cam[0], cam[1], cam[2], cam[3], cam[4]);
  if (!wpt_global_ctx()) die("init");
  // handleStoreMesh: parse the OBJ like the client (obj_parser.ts + the x8 transform), then the three-step upload
  if (!obj.empty()) {
    std::ifstream f(obj, std::ios::binary);
    if (!f) { std::fprintf(stderr, "wpt_render: cannot open %s\n", obj.c_str()); return 1; }
    std::stringstream ss; ss << f.rdbuf();
    std::string text = ss.str();
    int64_t nfloats = wpt_parse_obj(text.data(), text.size(), 1, nullptr, 0);
    if (nfloats < 0) die("parse_obj");
    std::vector<float> verts((size_t)nfloats);
    if (wpt_parse_obj(text.data(), text.size(), 1, verts.data(), verts.size()) != nfloats) die("parse_obj");
    const uint32_t mesh_id = 1;   // MESH_BUNNY_HIGH, scenes.rs:12
    wpt_allocate_mesh(mesh_id, (uint32_t)(nfloats / 3));
    float* dst = wpt_mesh_vertices(mesh_id);
    if (!dst) die("mesh_vertices");
    std::memcpy(dst, verts.data(), verts.size() * sizeof(float));
    if (wpt_notify_mesh_loaded(mesh_id) < 0) die("notify_mesh_loaded");
    if (!quiet) std::fprintf(stderr, "mesh: %lld triangles\n", (long long)(nfloats / 9));
  }
  wpt_update_settings(left, right, left_ad, right_ad, light_debug);
  if (*wpt_last_error()) die("update_settings");

  // run(): ticks of ~tick_ms each (worker.ts:76-84)
  using clock = std::chrono::steady_clock;
  uint64_t rays_per_tick = 1000, ticks = 0, samples = 0;
  const auto t_begin = clock::now();
  const uint8_t* frame = nullptr;
  for (;;) {
    double elapsed = std::chrono::duration<double>(clock::now() - t_begin).count();
    if ((max_ticks && ticks >= max_ticks) || (max_samples && samples >= max_samples) || (!max_ticks && !max_samples && elapsed >= seconds)) break;
    uint64_t n = rays_per_tick;
    if (max_samples && samples + n > max_samples) n = max_samples - samples;
    auto t0 = clock::now();
    wpt_compute(n);
    frame = wpt_results(sampling_view ? 1 : 0);   // synchronises, like reading the module's memory after compute()
    if (!frame) die("compute / results");
    double ms = std::chrono::duration<double, std::milli>(clock::now() - t0).count();
    samples += n; ticks += 1;
    if (ms <= 0.0) rays_per_tick = 1000;
    else { double scaled = (double)rays_per_tick * (tick_ms / ms); rays_per_tick = scaled < 1.0 ? 1 : (uint64_t)scaled; }
    if (!quiet) std::fprintf(stderr, "tick %llu: %llu samples in %.1f ms -> next %llu\n", (unsigned long long)ticks, (unsigned long long)n, ms, (unsigned long long)rays_per_tick);
  }
  if (!frame) frame = wpt_results(sampling_view ? 1 : 0);
  if (!frame) die("results");
  double total = std::chrono::duration<double>(clock::now() - t_begin).count();
  uint64_t st[8] = {0};
  wpt_ctx_stats(wpt_global_ctx(), st);
  std::printf("{\"ticks\": %llu, \"samples\": %llu, \"seconds\": %.3f, \"rays\": %llu, \"paths\": %llu, \"node_visits\": %llu, \"photons\": %llu, \"Mrays_per_s\": %.1f, \"out\": \"%s\"}\n",
              (unsigned long long)ticks, (unsigned long long)samples, total, (unsigned long long)st[0], (unsigned long long)st[1], (unsigned long long)st[2],
              (unsigned long long)st[4], total > 0 ? st[0] / total / 1e6 : 0.0, out.c_str());
  bool ok = out.size() > 4 && out.substr(out.size() - 4) == ".ppm" ? write_ppm(out, frame, w, h) : write_png(out, frame, w, h);
  if (!ok) { std::fprintf(stderr, "wpt_render: cannot write %s\n", out.c_str()); return 1; }
  return 0;
}
