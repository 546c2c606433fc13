"""One process per GPU: pixel partition + the glue between torch.distributed and the library's native NCCL plane.

The reference splits work by pixels (README.md:87: WebWorkers with pixel subsets; today a left/right viewport split,
src/wasm_interface.rs:78). Here the region's rows are cut into bands of 4 rows and rank r of `world` renders the bands b
with b % world == r (whole 8x4 warp tiles on every rank; csrc/wpt_types.h band_rows / band_row). Paths are independent
given the scene replica, so the only data-path exchanges are the accumulator all-gather and the photon-batch allreduce —
both issued by libwpt itself with NCCL on the session's stream (csrc/dist_nccl.cpp). torch.distributed is only the
rendezvous (`attach` broadcasts the NCCL unique id) and the barrier / timing plumbing of bench.py.

`exchange_rows` is the same exchange written with torch collectives: it runs on CPU tensors over gloo and is what
tests/test_dist_gloo.py uses to check the partition logic without a GPU.
"""
import numpy as np
import torch
import torch.distributed as dist

BAND = 4   # rows per band, WPT_BAND in csrc/wpt_types.h


def rows_of_rank(height, rank, world, y0=0):
    """Viewport rows owned by `rank`: the 4-row bands b of the region with b % world == rank."""
    return [y0 + y for y in range(height) if (y // BAND) % world == rank]


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, shape, dtype=torch.float32):
    """A torch view of library-owned device memory (PathTracer.device_buffers(), the reduce callback's pointer)."""
    typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.uint32: "<u4", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device="cuda")


_maps = {}


def _row_maps(height, world, device):
    """Per rank: the rows it owns (padded with -1 to the same length), as an index tensor (world, per)."""
    key = (height, world, str(device))
    if key not in _maps:
        rows = [rows_of_rank(height, r, world) for r in range(world)]
        per = max(len(x) for x in rows)
        idx = np.full((world, per), -1, np.int64)
        for r, x in enumerate(rows):
            idx[r, : len(x)] = x
        _maps[key] = torch.from_numpy(idx).to(device)
    return _maps[key]


def exchange_rows(frame, rank, world, group=None, y0=0, height=None):
    """All-gather the band-partitioned rows y0 .. y0 + height of `frame` ((H, ...) tensor, CPU or CUDA) in place.

    On entry rank r holds valid data in its own rows of the region; on exit every rank holds all of them.
    Three operations: pack own rows, one all_gather_into_tensor, one permuted unpack."""
    if world == 1:
        return frame
    height = frame.shape[0] - y0 if height is None else height
    region = frame[y0 : y0 + height]
    idx = _row_maps(height, world, frame.device)
    per = idx.shape[1]
    rest = tuple(frame.shape[1:])
    send = torch.zeros((per,) + rest, dtype=frame.dtype, device=frame.device)
    mine = idx[rank][idx[rank] >= 0]
    send[: mine.numel()] = region.index_select(0, mine)
    recv = torch.empty((world * per,) + rest, dtype=frame.dtype, device=frame.device)
    dist.all_gather_into_tensor(recv, send, group=group)   # concatenated along dim 0 (nccl and gloo)
    flat = idx.reshape(-1)
    valid = flat >= 0
    region.index_copy_(0, flat[valid], recv[valid])
    return frame


def broadcast_bytes(data, n, src=0, group=None):
    """Broadcast `n` bytes from rank `src` over the process group (works on nccl and gloo groups)."""
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.zeros(n, dtype=torch.uint8, device=device)
    if dist.get_rank(group) == src:
        t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().numpy().tobytes())


def attach(pt, rank, world, group=None):
    """Wire a PathTracer session into the job: band partition + the library's native NCCL plane (accumulator
    all-gather between adaptive rounds / on gather_frame(), photon batches merged with an integer sum-allreduce —
    every per-shot slot is written by exactly one rank, so the result is bit-identical to a single-GPU warm-up).
    torch.distributed only carries the 128-byte NCCL unique id from rank 0 to the others."""
    from .api import nccl_unique_id
    if world == 1:
        pt.attach_nccl(None, 0, 1)
        return pt
    uid = nccl_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128, 0, group)
    pt.attach_nccl(uid, rank, world)
    return pt
