/* wpt.h — C ABI of libwpt.so, the B200-native path-tracing core.
 *
 * Drop-in boundary for the reference's WASM export surface
 * (sourcedennis/wasm-pathtracer, src/wasm_interface.rs). Every `wpt_<name>` below replaces
 * the `#[wasm_bindgen] pub fn <name>` cited next to it: same argument order, same meaning,
 * primitives and raw pointers only (wasm_interface.rs:19-24). Buffers are owned by the
 * library; callers only borrow pointers (HOST memory).
 *
 * Errors: the reference panics (WASM trap). This ABI never unwinds: a failing call is a
 * no-op that records a message readable through wpt_last_error(); with the environment
 * variable WPT_STRICT=1 it abort()s instead, like the trap. There is no CPU fallback: every
 * compute entry point fails if no CUDA device is usable.
 *
 * Two API levels:
 *   1. the global-instance functions `wpt_init`, `wpt_compute`, ... — one implicit session,
 *      exactly like the reference's `static mut CONFIG` (wasm_interface.rs:62);
 *   2. handle functions `wpt_ctx_*` — the same operations on an explicit session, so that
 *      several sessions (one per GPU) can live in one process.
 */
#ifndef WPT_H
#define WPT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wpt_ctx wpt_ctx;

/* RenderType magic numbers, wasm_interface.rs:207-214 / PanelSettings.elm:94-99 */
enum { WPT_NO_NEE = 0, WPT_NORMAL_NEE = 1, WPT_PNEE = 2 };
/* Scene ids, wasm_interface.rs:389-398 / PanelScenes.elm:39-43 */
enum { WPT_SCENE_MUSEUM = 0, WPT_SCENE_BUNNY = 2 };
/* Extension, not a reference scene id (the reference panics on it): the commented-out Whitted scene of
 * scenes.rs:113-130 — textured Square floor (if texture 0 is loaded), refracting + reflecting Sphere. DESIGN.md 9. */
enum { WPT_SCENE_EXT_WHITTED = 256 };

/* ------------------------------------------------------------------ 1. global instance
 * (reference: `static mut CONFIG`, wasm_interface.rs:37-62) */

/* wasm_interface.rs:67-113  pub fn init(width,height,scene_id,cam_x,cam_y,cam_z,cam_rot_x,cam_rot_y)
 * A second call is an error ("Cannot init again", :74-76). Default settings as the
 * reference: left half NormalNEE + random strategy, right half PNEE + adaptive. */
void wpt_init(uint32_t width, uint32_t height, uint32_t scene_id,
              float cam_x, float cam_y, float cam_z, float cam_rot_x, float cam_rot_y);
/* wasm_interface.rs:120-134  pub fn results(is_show_sampling) -> *const u8
 * width*height*4 bytes RGBA8, row-major, top row first, A = 255. Borrowed HOST pointer,
 * valid until the next update_viewport. Synchronises outstanding GPU work. */
const uint8_t* wpt_results(uint32_t is_show_sampling);
/* wasm_interface.rs:137-148  pub fn reset() (public but not exported by the reference) */
void wpt_reset(void);
/* wasm_interface.rs:154-168  pub fn update_scene(scene_id) — ids 0 and 2 only */
void wpt_update_scene(uint32_t scene_id);
/* wasm_interface.rs:173-204  pub fn update_settings(left_type,right_type,is_left_adaptive,is_right_adaptive,is_light_debug) */
void wpt_update_settings(uint32_t left_type, uint32_t right_type, uint32_t is_left_adaptive,
                         uint32_t is_right_adaptive, uint32_t is_light_debug);
/* wasm_interface.rs:219-232  pub fn update_viewport(width,height) */
void wpt_update_viewport(uint32_t width, uint32_t height);
/* wasm_interface.rs:239-248  pub fn update_camera(cam_x,cam_y,cam_z,cam_rot_x,cam_rot_y) */
void wpt_update_camera(float cam_x, float cam_y, float cam_z, float cam_rot_x, float cam_rot_y);
/* wasm_interface.rs:259-270  pub fn allocate_mesh(id,num_vertices) */
void wpt_allocate_mesh(uint32_t id, uint32_t num_vertices);
/* wasm_interface.rs:275-288  pub fn mesh_vertices(id) -> *mut Vec3 — num_vertices*3 packed f32 */
float* wpt_mesh_vertices(uint32_t id);
/* wasm_interface.rs:293-329  pub fn notify_mesh_loaded(id) -> bool */
int wpt_notify_mesh_loaded(uint32_t id);
/* wasm_interface.rs:335-352  pub fn allocate_texture(id,width,height) -> *mut (u8,u8,u8) */
uint8_t* wpt_allocate_texture(uint32_t id, uint32_t width, uint32_t height);
/* wasm_interface.rs:357-366  pub fn notify_texture_loaded(id) -> bool (always false) */
int wpt_notify_texture_loaded(uint32_t id);
/* wasm_interface.rs:374-384  pub fn compute(num_samples) — asynchronous on the device;
 * wpt_results / wpt_stats synchronise. */
void wpt_compute(uint64_t num_samples);

/* Message of the most recent failure on this thread ("" if none). */
const char* wpt_last_error(void);
/* The global session as a handle (NULL before wpt_init). */
wpt_ctx* wpt_global_ctx(void);

/* ------------------------------------------------------------------ 2. handle API */

/* Create a session on CUDA device `device` (-1 = current device). Returns NULL on failure.
 * WPT_DEVICE_NONE makes a host-only session: scene setup, BVH build / collapse and mesh
 * loading work and can be inspected (wpt_ctx_scene_info, wpt_ctx_bvh2, ...), every call that
 * would compute fails — there is no CPU rendering path. */
#define WPT_DEVICE_NONE (-2)
wpt_ctx* wpt_ctx_create(int device, uint32_t width, uint32_t height, uint32_t scene_id,
                        float cam_x, float cam_y, float cam_z, float cam_rot_x, float cam_rot_y);
void wpt_ctx_destroy(wpt_ctx* ctx);
/* Same operations as the global functions; return 0 on success, -1 on error. */
const uint8_t* wpt_ctx_results(wpt_ctx* ctx, uint32_t is_show_sampling);
int wpt_ctx_reset(wpt_ctx* ctx);
int wpt_ctx_update_scene(wpt_ctx* ctx, uint32_t scene_id);
int wpt_ctx_update_settings(wpt_ctx* ctx, uint32_t left_type, uint32_t right_type, uint32_t is_left_adaptive,
                            uint32_t is_right_adaptive, uint32_t is_light_debug);
int wpt_ctx_update_viewport(wpt_ctx* ctx, uint32_t width, uint32_t height);
int wpt_ctx_update_camera(wpt_ctx* ctx, float cam_x, float cam_y, float cam_z, float cam_rot_x, float cam_rot_y);
int wpt_ctx_allocate_mesh(wpt_ctx* ctx, uint32_t id, uint32_t num_vertices);
float* wpt_ctx_mesh_vertices(wpt_ctx* ctx, uint32_t id);
int wpt_ctx_notify_mesh_loaded(wpt_ctx* ctx, uint32_t id);   /* 1 / 0 = the reference's bool, -1 = error */
uint8_t* wpt_ctx_allocate_texture(wpt_ctx* ctx, uint32_t id, uint32_t width, uint32_t height);
int wpt_ctx_notify_texture_loaded(wpt_ctx* ctx, uint32_t id);
int wpt_ctx_compute(wpt_ctx* ctx, uint64_t num_samples);

/* ------------------------------------------------------------------ 3. additions
 * Not in the reference; needed to express the benchmark configurations (BVH4, full-frame
 * exact-spp rendering, multi-GPU row partitions) and to export what the reference keeps
 * internal (`num_bvh_hits`, tracer.rs:40 — documented as returned by compute but never
 * exported, wasm_interface.rs:371-374). */

typedef struct wpt_config {
  uint32_t bvh_kind;        /* 2 (reference default, scene.rs:60) or 4 (bvh4.rs collapse)          */
  uint32_t render_type;     /* WPT_NO_NEE / WPT_NORMAL_NEE / WPT_PNEE for the exact/adaptive drivers */
  uint32_t light_debug;     /* tracer.rs:46-49 is_debug_photons                                      */
  uint32_t base_seed;       /* base of the per-path stream hash (default 0xBABABEBE, rng.rs:11)      */
  uint64_t photon_target;   /* diffuse-hit photons to collect (default 300000, tracer.rs:104)        */
  uint32_t region_x, region_y, region_w, region_h; /* logical sampling region (0 size = full frame)  */
  uint32_t rank, world;     /* this session renders rows {y : band(y) == rank} of the region         */
  uint32_t engine;          /* 0 = persistent path kernel k_mega (default), 1 = multi-kernel wavefront, 4 = experimental warp-pool kernel k_wpool */
  uint32_t reserved[3];
} wpt_config;

void wpt_default_config(wpt_config* cfg);
int wpt_ctx_set_config(wpt_ctx* ctx, const wpt_config* cfg);
int wpt_ctx_get_config(wpt_ctx* ctx, wpt_config* cfg);

/* `spp` more samples for every pixel of this session's rows, exact counts (mode-B driver). */
int wpt_ctx_render_exact(wpt_ctx* ctx, uint32_t spp);
/* Adaptive sampling (sampling_strategy.rs:77-220) over the region with a tick budget;
 * returns the ticks consumed over the whole region (all ranks), -1 on error. */
int64_t wpt_ctx_render_adaptive(wpt_ctx* ctx, uint64_t budget_ticks);
/* Random strategy (sampling_strategy.rs:30-71): `ticks` uniform pixel picks with replacement. */
int wpt_ctx_render_random(wpt_ctx* ctx, uint64_t ticks);
/* Multi-GPU: called between adaptive rounds so that the caller can exchange the accumulator
 * rows of the other ranks (the next error map reads the whole region). NULL removes it. */
int wpt_ctx_set_exchange_callback(wpt_ctx* ctx, void (*callback)(void* user), void* user);
/* Multi-GPU photon warm-up (tracer.rs:126-152 split over ranks): with this callback set and
 * config.world > 1, rank r emits the shots r, r + world, ... of every batch into per-shot slots
 * and calls back to have `n_words` 32-bit words at device pointer `dev_words` summed in place
 * over all ranks (integer sum: every slot is written by one rank, so the merge is bit-exact),
 * on the session's stream (e.g. ncclAllReduce(ncclUint32, ncclSum)). NULL removes it: every
 * rank then emits all shots itself. */
int wpt_ctx_set_reduce_callback(wpt_ctx* ctx, void (*callback)(void* user, void* dev_words, uint64_t n_words), void* user);
/* Native multi-GPU plane: NCCL inside the library — replaces the reference's only cross-worker data hand-off,
 * the SharedArrayBuffer copy of src_ts/worker/worker.ts:84-89 (and the old 8-worker pixel split, README.md:87).
 * One process (or thread) per GPU. Rank 0 calls wpt_nccl_unique_id and the host distributes the 128 bytes by any means
 * (MPI, a file, torch.distributed); every rank then attaches its session: ncclCommInitRank, config.rank / world, and the
 * built-in exchange (accumulator all-gather of the region's rows, between adaptive rounds and on wpt_ctx_gather_frame)
 * and photon-batch reduction (ncclAllReduce of uint32, sum) replace the two callbacks above. All collectives are issued
 * on the session's stream. NCCL is dlopen()ed at attach time (libnccl.so.2). `_comm` borrows an existing ncclComm_t. */
int wpt_nccl_unique_id(uint8_t out[128]);
int wpt_ctx_attach_nccl(wpt_ctx* ctx, const uint8_t id[128], uint32_t rank, uint32_t world);
int wpt_ctx_attach_nccl_comm(wpt_ctx* ctx, void* nccl_comm, uint32_t rank, uint32_t world);
int wpt_ctx_detach_nccl(wpt_ctx* ctx);
/* All-gather the accumulators of the region's rows over the attached ranks (every rank ends up with the whole region). */
int wpt_ctx_gather_frame(wpt_ctx* ctx);
/* Photon warm-up (tracer.rs:103-152) + octree light-CDF build (photon_tree.rs). */
int wpt_ctx_build_photons(wpt_ctx* ctx);
/* Block until queued GPU work is finished. */
int wpt_ctx_synchronize(wpt_ctx* ctx);

/* Counters since the last reset: out[0]=rays (trace_g calls), [1]=paths, [2]=BVH node
 * visits (the reference's num_bvh_hits), [3]=photons shot, [4]=photons stored,
 * [5]=wavefront iterations, [6]=kernel launches, [7]=reserved. */
int wpt_ctx_stats(wpt_ctx* ctx, uint64_t out[8]);
/* Primary-ray probe of sample 0 of every viewport pixel: hit shape index (-1 = miss),
 * node-visit count and hit distance. Any output pointer may be NULL. */
int wpt_ctx_primary_probe(wpt_ctx* ctx, int32_t* ids, uint32_t* visits, float* dist);
/* Accumulators (render_target.rs:8-9): rgb = width*height*3 f32 sums, counts = samples. */
int wpt_ctx_accum(wpt_ctx* ctx, float* rgb, uint32_t* counts);
/* Trace a batch of rays through Scene::trace_g (scene.rs:162-184) on the device. */
int wpt_ctx_trace_rays(wpt_ctx* ctx, const float* origins, const float* dirs, uint64_t n,
                       int32_t* ids, float* dist, uint32_t* visits, float* normals);

/* Scene introspection: info[0]=shapes, [1]=infinite shapes, [2]=lights, [3]=BVH2 node array
 * length, [4]=BVH2 depth, [5]=BVH4 node array length, [6]=bvh kind, [7]=BVH4 depth. */
int wpt_ctx_scene_info(wpt_ctx* ctx, uint64_t info[8]);
int wpt_ctx_bvh2(wpt_ctx* ctx, float* bounds6, uint32_t* left_first, uint32_t* count);
int wpt_ctx_bvh4(wpt_ctx* ctx, float* bounds24, int32_t* children4, uint32_t* num_children);
int wpt_ctx_shape_order(wpt_ctx* ctx, int32_t* source_index, int32_t* type);
int wpt_ctx_lights(wpt_ctx* ctx, uint32_t* shape_index);

/* Photon read-backs (after wpt_ctx_build_photons). */
int64_t wpt_ctx_photon_count(wpt_ctx* ctx, uint64_t* shots);
int wpt_ctx_photon_list(wpt_ctx* ctx, uint32_t* light, float* loc3, float* weight);
/* Octree in DFS pre-order: meta[n*3] = depth, is_node, photons in cell; cum/bins [n*L].
 * Call with NULL buffers for the node count. */
int64_t wpt_ctx_photon_tree(wpt_ctx* ctx, uint32_t* meta, float* cum, float* bins);
int wpt_ctx_photon_sample(wpt_ctx* ctx, const float* pts3, const uint32_t* seeds, uint64_t n,
                          uint32_t* light, float* pdf);
/* Adaptive read-backs: error map of the region + {min, avg, max}; last round's spp. */
int wpt_ctx_error_map(wpt_ctx* ctx, float* mse, float stats3[3]);
int wpt_ctx_round_spp(wpt_ctx* ctx, uint32_t* spp);

/* Use an externally owned CUDA stream (e.g. torch's current stream) for all GPU work of
 * this session; 0 restores the session's own stream. */
int wpt_ctx_set_stream(wpt_ctx* ctx, uint64_t cuda_stream);
/* Re-send the flattened scene (nodes, shapes, materials, lights) from pinned host memory to
 * the device; returns the bytes copied. Lets a caller time host->device traffic honestly. */
int64_t wpt_ctx_upload_scene(wpt_ctx* ctx);
/* Per-launch CUDA-event timing of the wavefront kernels. After wpt_ctx_profile(ctx, 1):
 * out[0]=trace-kernel ms, [1]=trace launches, [2]=shade-kernel ms, [3]=shade launches,
 * [4]=leaf primitive tests, [5]=rays, [6]=node visits, [7]=reserved — accumulated since the
 * last wpt_ctx_profile call. */
int wpt_ctx_profile(wpt_ctx* ctx, int enable);
int wpt_ctx_profile_read(wpt_ctx* ctx, double out[8]);
/* Strategy rounds since wpt_ctx_profile(ctx, 1): out[0] = adaptive rounds, [1] = error-map ms, [2] = round render ms, [3] = exchange ms
 * (CUDA events on the session's stream), [4] = wall-clock ms of the last photon warm-up, [5] = NCCL collectives issued. */
int wpt_ctx_profile_read_rounds(wpt_ctx* ctx, double out[8]);

/* Device pointers for zero-copy collectives (multi-GPU plumbing lives above this ABI). */
int wpt_ctx_device_buffers(wpt_ctx* ctx, uint64_t ptrs[8], uint64_t sizes[8]);
/* After an external exchange wrote other ranks' rows into the accumulators / RGBA buffer. */
int wpt_ctx_mark_accum_dirty(wpt_ctx* ctx);

/* OBJ file -> mesh `id`, reproducing src_ts/client/obj_parser.ts:3-51 and the client
 * transform (x8, x8, x-8) of src_ts/client/index.ts:216-220. Returns vertices loaded. */
int64_t wpt_ctx_load_obj(wpt_ctx* ctx, uint32_t id, const char* path, int apply_client_scale);
int64_t wpt_parse_obj(const char* text, uint64_t len, int apply_client_scale, float* out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* WPT_H */
