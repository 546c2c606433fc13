// The multi-GPU data plane inside libwpt: NCCL all-gather of the accumulator rows and the integer
// sum-allreduce of the photon batches, issued on the session's stream. It replaces the only cross-worker data transfer of
// the reference, the SharedArrayBuffer hand-off of src_ts/worker/worker.ts:84-89 (and its old 8-worker pixel split,
// README.md:87). NCCL is bound with dlopen at attach time: libwpt.so has no link-time dependency on it, and inside a
// process that already loaded a libnccl.so.2 (torch) the same copy is used.
#include "context.h"
#include <dlfcn.h>
#include <nccl.h>   // types and enums only

namespace wpt {

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl() {
  static NcclApi api;
  if (api.lib) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
  if (!api.lib) throw std::runtime_error(std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror());
  auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) throw std::runtime_error(std::string("NCCL symbol missing: ") + n); return p; };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  return api;
}
void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error: ") + nccl().GetErrorString(r) + " in " + what);
}
#define WPT_NCCL(x) nccl_check((x), #x)
}  // namespace

void nccl_unique_id(uint8_t out[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  WPT_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(out, &id, 128);
}

void Context::attach_nccl(const uint8_t id128[128], void* existing_comm, uint32_t rank, uint32_t world) {
  require_device();
  if (world == 0 || rank >= world) throw std::runtime_error("invalid rank/world");
  detach_nccl();
  WPT_CUDA(cudaSetDevice(device));
  if (existing_comm) { nccl_comm = existing_comm; nccl_own = false; nccl(); }
  else if (world > 1) {
    ncclUniqueId id; std::memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    WPT_NCCL(nccl().CommInitRank(&comm, (int)world, id, (int)rank));
    nccl_comm = comm; nccl_own = true;
  }
  cfg.rank = rank; cfg.world = world;
  slots = 0;   // the pixel map depends on the partition
  if (nccl_comm && world > 1) {   // NCCL sets up its channels at the first collective (~1 s): pay that here, not inside the first render
    WPT_CUDA(cudaMemsetAsync(w_work.p + 2, 0, sizeof(uint32_t), stream));
    WPT_NCCL(nccl().AllReduce(w_work.p + 2, w_work.p + 2, 1, ncclUint32, ncclSum, static_cast<ncclComm_t>(nccl_comm), stream));
    // ... and NCCL picks its kernels by message size and loads them lazily (first photon warm-up on 8 GPUs: 113 ms instead of 6):
    // one collective of each kind at the sizes the plane uses — a photon batch (131 072 shots x 6 words) and a band of accumulator rows
    {
      DevBuf<uint32_t> warm, gath;
      const size_t words = 131072u * 6u, part = 1u << 20;
      warm.alloc(std::max(words, part)); gath.alloc(part * world);
      WPT_CUDA(cudaMemsetAsync(warm.p, 0, std::max(words, part) * sizeof(uint32_t), stream));
      WPT_NCCL(nccl().AllReduce(warm.p, warm.p, words, ncclUint32, ncclSum, static_cast<ncclComm_t>(nccl_comm), stream));
      WPT_NCCL(nccl().AllGather(warm.p, gath.p, part, ncclFloat, static_cast<ncclComm_t>(nccl_comm), stream));
      WPT_CUDA(cudaStreamSynchronize(stream));
    }
    WPT_CUDA(cudaStreamSynchronize(stream));
  }
  if (world > 1) {
    exchange_hook = [this] { exchange_native(); };
    reduce_hook = [this](uint32_t* p, uint64_t n) { reduce_native(p, n); };
  } else { exchange_hook = nullptr; reduce_hook = nullptr; }
}

void Context::detach_nccl() {
  if (nccl_comm && nccl_own && has_device) { cudaStreamSynchronize(stream); nccl().CommDestroy(static_cast<ncclComm_t>(nccl_comm)); }
  if (nccl_comm) { exchange_hook = nullptr; reduce_hook = nullptr; }
  nccl_comm = nullptr; nccl_own = false;
}

// All-gather of the accumulator rows of the current region: pack own rows, ncclAllGather, permuted unpack — three
// operations on the session's stream, no host synchronisation. Every rank ends up with every row of the region.
void Context::exchange_native() {
  require_device();
  if (!nccl_comm || cfg.world <= 1) return;
  uint32_t rx, ry, rw, rh;
  region(&rx, &ry, &rw, &rh);
  const uint32_t world = cfg.world, rank = cfg.rank;
  uint32_t per = 0;
  for (uint32_t r = 0; r < world; r++) per = std::max(per, band_rows(rh, r, world));
  if (!per || !rw) return;
  const size_t chunk = (size_t)per * rw;   // float4 per rank
  x_send.alloc(chunk); x_recv.alloc(chunk * world);
  launch_pack_rows(d_accum.p, W, rx, ry, rw, rh, rank, world, per, x_send.p, stream);
  WPT_NCCL(nccl().AllGather(x_send.p, x_recv.p, chunk * 4, ncclFloat, static_cast<ncclComm_t>(nccl_comm), stream));
  launch_unpack_rows(x_recv.p, W, rx, ry, rw, rh, rank, world, per, d_accum.p, stream);
  launches += 2; collectives += 1;
  rgba_stale = true;
}

void Context::reduce_native(uint32_t* dev_words, uint64_t n) {
  require_device();
  if (!nccl_comm || cfg.world <= 1 || !n) return;
  WPT_NCCL(nccl().AllReduce(dev_words, dev_words, n, ncclUint32, ncclSum, static_cast<ncclComm_t>(nccl_comm), stream));
  collectives += 1;
}

}  // namespace wpt
