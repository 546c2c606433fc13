"""world_size-2 / -3 gloo tests (CPU) of the multi-GPU host logic: band partition, framebuffer exchange, id broadcast."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, H, W_, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wasm_pathtracer_b200.dist import exchange_rows, rows_of_rank
    frame = torch.zeros((H, W_, 4), dtype=torch.float32)
    for y in rows_of_rank(H, rank, world):
        frame[y] = float(rank + 1) * 1000 + y
    exchange_rows(frame, rank, world)
    q.put((rank, frame.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, H):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, 5, q)) for r in range(world)]
    for p in procs: p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs: p.join(60)
    want = np.zeros((H, 5, 4), np.float32)
    for y in range(H):
        want[y] = float((y // 4) % world + 1) * 1000 + y      # 4-row bands, band b belongs to rank b % world
    for r in range(world):
        assert np.array_equal(res[r], want)


def test_exchange_rows_even():
    _run(2, 16)


def test_exchange_rows_ragged():
    _run(2, 7)     # one full band for rank 0, a ragged 3-row band for rank 1
    _run(3, 10)    # rank 2 owns only the ragged last band; and a world that does not divide the band count


def test_exchange_rows_more_ranks_than_bands():
    _run(3, 5)     # rank 2 owns nothing


def test_rows_of_rank_cover_the_region():
    from wasm_pathtracer_b200.dist import rows_of_rank
    for H in (1, 7, 1080):
        for world in (1, 2, 3, 8):
            rows = sorted(sum((rows_of_rank(H, r, world, 10) for r in range(world)), []))
            assert rows == list(range(10, 10 + H))


def _reduce_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wasm_pathtracer_b200.dist import broadcast_bytes
    uid = broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)     # the NCCL unique id travels like this
    assert uid == bytes(range(128))

    def allreduce_words(t):   # what ncclAllReduce(ncclUint32, ncclSum) does inside libwpt (csrc/dist_nccl.cpp)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    # the photon-batch merge: slot i is written by rank i % world only (raw float bits viewed as int32), zeros elsewhere
    n = 1000
    rng = np.random.default_rng(7)
    full = rng.standard_normal(n).astype(np.float32).view(np.int32)
    mine = np.where(np.arange(n) % world == rank, full, 0).astype(np.int32)
    t = torch.from_numpy(mine.copy())
    allreduce_words(t)
    q.put((rank, t.numpy().copy(), full))
    dist.barrier()
    dist.destroy_process_group()


def test_photon_slot_merge_is_an_exact_integer_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_reduce_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(60)
    for rank, got, full in res:
        assert np.array_equal(got, full)      # bit-exact merge: x + 0 == x in integer arithmetic
