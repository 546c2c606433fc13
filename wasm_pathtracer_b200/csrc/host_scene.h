// Host-side scene setup, BVH2 build, BVH4 collapse and flattening into device records.
// Reference behaviour (what must come out) is cited per function; the implementation is
// index-based and in place (no per-level shape copies), see DESIGN.md "Host builder".
#pragma once
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>
#include "wpt_types.h"

namespace wpt {

struct V3 { float x, y, z; };
struct Box { float lo[3], hi[3]; };

struct HostShape {
  ShapeType type;
  uint32_t mat;       // index into HostScene::mats
  float p[9];         // triangle: v0,v1,v2 | plane: location, normal | torus: location, R, r | aa_rect: lo, hi | sphere: location, radius | square: location, size
  int32_t source;     // index in the scene's original shape list
};
struct HostMaterial { float r, g, b; bool emissive; uint32_t kind = MAT_DIFFUSE; float param = 0.0f; uint32_t tex = 0; };

struct HostBVH2Node { Box box; uint32_t left_first, count; };
struct HostBVH4Node { Box child[4]; int32_t children[4]; uint32_t num_children; };

struct HostScene {
  float bg[3] = {0, 0, 0};
  std::vector<HostShape> shapes;        // final order: infinite shapes, then BVH order (bvh.rs:103-125)
  std::vector<HostMaterial> mats;
  std::vector<uint32_t> lights;         // shape indices of emissive shapes, ascending (scene.rs:62-66)
  uint32_t num_inf = 0;
  std::vector<HostBVH2Node> bvh2;
  std::vector<HostBVH4Node> bvh4;       // empty unless bvh_kind == 4
  uint32_t bvh_kind = 2;
  uint32_t depth2 = 0, depth4 = 0;
};

// scenes.rs:15-68 / :75-111
void scene_museum(std::vector<HostShape>& shapes, std::vector<HostMaterial>& mats);
void scene_bunny(const std::vector<HostShape>* mesh, std::vector<HostShape>& shapes, std::vector<HostMaterial>& mats);
// extension scene 256: the commented-out Whitted scene of scenes.rs:113-130; `with_floor` = texture 0 is loaded
void scene_whitted(bool with_floor, std::vector<HostShape>& shapes, std::vector<HostMaterial>& mats, float bg[3]);
// wasm_interface.rs:297-313 — `verts` = 9 floats per triangle; material slot `mat`
std::vector<HostShape> mesh_triangles(const float* verts, size_t num_vertices, uint32_t mat);
// scene.rs:43-69 (+ bvh.rs, bvh4.rs)
void build_scene(HostScene& sc, std::vector<HostShape> shapes, std::vector<HostMaterial> mats, uint32_t bvh_kind, uint32_t num_bins = 16);
// obj_parser.ts:3-51 (+ index.ts:216-220 when `client_scale`)
std::vector<float> parse_obj_text(const char* text, size_t len, bool client_scale);

// Flatten into device records (wpt_types.h)
void flatten_scene(const HostScene& sc, std::vector<DNode2>& n2, std::vector<DNode4>& n4, std::vector<DShape>& shp,
                   std::vector<DMaterial>& mats, std::vector<DLight>& lights);

}  // namespace wpt
