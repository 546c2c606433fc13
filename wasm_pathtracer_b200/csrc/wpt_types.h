// Device data layout shared by the host flattener and the CUDA kernels.
// Everything is laid out for 128-bit (__ldg float4) fetches; see DESIGN.md "Data layout".
#pragma once
#include <stdint.h>
#include <vector_types.h>

namespace wpt {

// src/math/mod.rs:11
#define WPT_EPSILON 0.0002f
#define WPT_PI 3.14159265358979323846f

enum ShapeType : uint32_t { SH_TRIANGLE = 0, SH_PLANE = 1, SH_TORUS = 2, SH_AARECT = 3, SH_SPHERE = 4, SH_SQUARE = 5 };
// Material kinds (DMaterial::c.w). Only DIFFUSE and EMISSIVE exist in the reference (material.rs:16-20); the others are the
// extension of DESIGN.md 9 (named by the task, removed from the reference: parity unpinned).
enum MatKind : uint32_t { MAT_DIFFUSE = 0, MAT_EMISSIVE = 1, MAT_REFLECT = 2, MAT_REFRACT = 3, MAT_DIFFUSE_TEX = 4 };

// BVH2 node, 32 B like the reference's BVHNode (bvh.rs:14-20); siblings are adjacent so the
// two child boxes tested at scene.rs:242-243 are one 64 B fetch.
//   a = (x_min, y_min, z_min, x_max)   b = (y_max, z_max, bits(left_first), bits(count))
struct DNode2 { float4 a, b; };

// BVH4 node, 128 B = one cache line like BVHNode4 (bvh4.rs:17-26), SoA lanes.
struct DNode4 {
  float4 x_min, y_min, z_min, x_max, y_max, z_max;
  int4 children;          // >= 0 inner node index; < 0 leaf code (bit31, count<<27, first)
  uint32_t num_children, pad0, pad1, pad2;
};

// Shape record, 64 B. meta = type | (material index << 8)
//   triangle: q0=(v0, meta) q1=(v1, n.x) q2=(v2, n.y) q3=(n.z, nn.x, nn.y, nn.z)
//             n = (v1-v0)x(v2-v0) un-normalised, nn = n.normalize()   (triangle.rs:164,182)
//   plane   : q0=(location, meta) q1=(normal, normal.location)        (plane.rs:80-99)
//   torus   : q0=(location, meta) q1=(big_r, small_r, 0, 0)           (torus.rs:11-16)
//   aa_rect : q0=(x_min,y_min,z_min, meta) q1=(x_max,y_max,z_max,0)   (aa_rect.rs:8-16)
//   sphere  : q0=(location, meta) q1=(radius, 0, 0, 0)                (sphere.rs:10-15)
//   square  : q0=(location, meta) q1=(size, 0, 0, 0)                  (square.rs:11-16)
struct DShape { float4 q0, q1, q2, q3; };

// Material (material.rs:16-20): c = (r,g,b, kind); r,g,b = colour, intensity (emissive) or absorption (refract);
// p.x = reflection share (reflect) / index of refraction (refract), p.y = bits(texture slot) (textured diffuse)
struct DMaterial { float4 c; float4 p; };

// Texture (texture.rs:9-13): packed RGB8, row-major
struct DTexture { const uint8_t* rgb; uint32_t width, height; };
#define WPT_MAX_TEXTURES 4

// Area light = emissive triangle (scene.rs:62-66): shape index, Heron area (triangle.rs:70-78),
// unit normal (triangle.rs:104) and intensity.
struct DLight {
  float4 n_area;      // normal.xyz, area
  float4 intensity;   // rgb, bits(shape index)
};

struct DCamera {
  float ox, oy, oz;
  float cx, sx, cy, sy;     // cos/sin of rot_x, rot_y — evaluated once on the host (vec3.rs:95-119)
  float w_inv, h_inv, ar;   // tracer.rs:163-172
};

// Flattened photon octree (photon_tree.rs). Node i: child_base[i] = index of its first child
// (8 children contiguous, octant order) or 0xFFFFFFFF for a leaf; cum[i*L .. i*L+L) = CDF.
struct DPhotonTree {
  const uint32_t* child_base;
  const uint32_t* nbr;      // [node][27]: neighbour cell per direction (dx+1) + 3(dy+1) + 9(dz+1), leaves only
  const float* cum;
  uint32_t num_lights;
  uint32_t num_nodes;
};

struct DScene {
  const DNode2* nodes2;
  const DNode4* nodes4;
  const DShape* shapes;
  const DMaterial* mats;
  const DLight* lights;
  uint32_t num_inf, num_shapes, num_lights, bvh_kind;
  float bg_r, bg_g, bg_b, pad;
  DTexture tex[WPT_MAX_TEXTURES];   // extension: textured diffuse
};

// Multi-GPU pixel partition: the region's rows are cut into bands of WPT_BAND rows and rank r of `world`
// owns the bands b with b % world == r — whole 8x4 warp tiles on every rank, and the bunny / sky imbalance still averaged out.
#define WPT_BAND 4u
#if defined(__CUDACC__)
#define WPT_HD __host__ __device__ inline
#else
#define WPT_HD inline
#endif
WPT_HD uint32_t band_rows(uint32_t h, uint32_t rank, uint32_t world) {   // rows of an h-row region owned by `rank`
  const uint32_t nb = (h + WPT_BAND - 1u) / WPT_BAND;
  const uint32_t mine = nb > rank ? (nb - rank + world - 1u) / world : 0u;
  if (!mine) return 0u;
  uint32_t rows = mine * WPT_BAND;
  if (rank + (mine - 1u) * world == nb - 1u) rows -= nb * WPT_BAND - h;   // the ragged last band
  return rows;
}
WPT_HD uint32_t band_row(uint32_t r, uint32_t rank, uint32_t world) {     // the rank's r-th row -> row of the region
  return ((r / WPT_BAND) * world + rank) * WPT_BAND + (r % WPT_BAND);
}

// Mode-B accumulation contract B10: render_exact sums a pixel's samples in segments of this length
#define WPT_SEGMENT_LEN 8u

// xorshift32 stream contract (DESIGN.md "RNG contract")
enum : uint32_t { STREAM_PATH = 1, STREAM_PHOTON = 2, STREAM_PIXEL = 3 };

}  // namespace wpt
