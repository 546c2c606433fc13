"""One process per GPU: row partition + framebuffer exchange over torch.distributed.

The reference splits work by pixels (README.md:87: WebWorkers with pixel subsets; today a
left/right viewport split, src/wasm_interface.rs:78). Here rank r of `world` renders the
rows y with (y - region_y) % world == r; paths are independent given the scene replica, so
the only data-path exchange is the accumulator all-gather below (plus the photon and
adaptive reductions in `DistributedPathTracer`). Collectives run on the session's stream.
"""
import numpy as np
import torch
import torch.distributed as dist


def rows_of_rank(height, rank, world, y0=0):
    """Viewport rows owned by `rank` (interleaved)."""
    return list(range(y0 + rank, y0 + height, world))


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, shape, dtype=torch.float32):
    typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.uint32: "<u4", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device="cuda")


_staging = {}


def exchange_rows(frame, rank, world, group=None):
    """All-gather interleaved rows of `frame` ((H, ...) tensor, CPU or CUDA) in place.

    On entry rank r holds valid data in rows r, r + world, ...; on exit every rank holds all rows.
    Three device operations: pack own rows, one all_gather_into_tensor, one permuted unpack.
    """
    if world == 1:
        return frame
    H = frame.shape[0]
    rest = tuple(frame.shape[1:])
    per = (H + world - 1) // world
    key = (H, rest, frame.dtype, str(frame.device), world)
    if key not in _staging:
        _staging[key] = (torch.zeros((per,) + rest, dtype=frame.dtype, device=frame.device),
                         torch.empty((world, per) + rest, dtype=frame.dtype, device=frame.device))
    send, recv = _staging[key]
    mine = frame[rank::world]
    send[: mine.shape[0]].copy_(mine)
    dist.all_gather_into_tensor(recv.view((world * per,) + rest), send, group=group)   # concatenated along dim 0 (nccl and gloo)
    full = H // world                       # rows every rank owns
    if full:
        frame[: full * world].view((full, world) + rest).copy_(recv[:, :full].transpose(0, 1))
    for r in range(H - full * world):       # ragged tail: ranks r < H mod world own one more row
        frame[full * world + r].copy_(recv[r, full])
    return frame


def session_stream(pt):
    """The CUDA stream the session issues its kernels on, as a torch stream."""
    return torch.cuda.ExternalStream(pt.device_buffers()["stream"])


def allgather_rows(pt, rank, world, group=None):
    """Exchange the accumulators (rgb sums + sample counts) of a PathTracer session.

    The collective is issued on the session's own stream, so it is ordered after the render
    kernels and before whatever the session launches next (no host synchronisation)."""
    ptr, _ = pt.device_buffers()["accum"]
    with torch.cuda.stream(session_stream(pt)):
        acc = device_tensor(ptr, (pt.H, pt.W, 4), torch.float32)
        exchange_rows(acc, rank, world, group)
    pt.mark_accum_dirty()
    return acc


def allreduce_words(words, group=None):
    """In-place integer sum of a 1-D int32 tensor over all ranks (NCCL on CUDA tensors, gloo on CPU)."""
    dist.all_reduce(words, op=dist.ReduceOp.SUM, group=group)
    return words


def attach(pt, rank, world, group=None):
    """Wire a PathTracer session into the process group: row partition, accumulator all-gather between
    adaptive rounds, and the photon warm-up split over ranks (each batch of per-shot photon slots is
    merged with an integer sum-allreduce on the session's stream — every slot is written by exactly
    one rank, so the result is bit-identical to a single-GPU warm-up)."""
    pt.set_config(rank=rank, world=world)
    if world == 1:
        pt.set_exchange_callback(None); pt.set_reduce_callback(None)
        return pt

    def reduce(ptr, n):
        with torch.cuda.stream(session_stream(pt)):
            allreduce_words(device_tensor(ptr, (n,), torch.int32), group)

    pt.set_exchange_callback(lambda: allgather_rows(pt, rank, world, group))
    pt.set_reduce_callback(reduce)
    return pt
