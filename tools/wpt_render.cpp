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
// Build: g++ -O2 -std=c++17 tools/wpt_render.cpp -Iinclude -Lwasm_pathtracer_b200 -lwpt -Wl,-rpath,'$ORIGIN/../wasm_pathtracer_b200' -o tools/wpt_render
#include <chrono>
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

int main(int argc, char** argv) {
  uint32_t scene = WPT_SCENE_BUNNY, w = 512, h = 512;
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
    else if (a == "--camera") { if (std::sscanf(next(), "%f,%f,%f,%f,%f", cam, cam + 1, cam + 2, cam + 3, cam + 4) != 5) { std::fprintf(stderr, "--camera x,y,z,rx,ry\n"); return 2; } cam_set = true; }
    else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }
  if (!cam_set && scene == WPT_SCENE_MUSEUM) { const float m[5] = {0.0f, 16.34f, -23.76f, 0.54f, 0.0f}; std::memcpy(cam, m, sizeof m); }   // index.ts:156

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
