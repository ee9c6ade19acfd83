"""SNP-sharded multi-GPU plumbing: one process per GPU, torch.distributed for
rendezvous and host-side collectives; the per-op shard sum itself runs inside
the native library (fpb_comm_init: a kernel over NVLink peer memory, csrc/fpb_peer.cuh,
or ncclAllReduce on the library's communicator when peer memory cannot be mapped).

X X' x = sum_g X_g X_g' x over disjoint contiguous SNP ranges -- the block sum
of svdwide.cpp:48-59, distributed."""
from __future__ import annotations

import ctypes

import numpy as np


def shard_range(nsnps: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous SNP range [j0, j1) of `rank`, balanced by count."""
    base, rem = divmod(nsnps, world)
    j0 = rank * base + min(rank, rem)
    return j0, j0 + base + (1 if rank < rem else 0)


def gather_meansd(local_meansd: np.ndarray, nsnps: int, world: int, rank: int) -> np.ndarray:
    """Concatenate per-shard (mean, sd) rows into the full nsnps x 2 table
    (Data::X_meansd is column-local, data.cpp:290-291)."""
    import torch
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(local_meansd))
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == nsnps
    return np.asfortranarray(out)


def attach_nccl(op, world: int, rank: int) -> None:
    """Create the library-side NCCL communicator: rank 0 makes the unique id,
    torch.distributed broadcasts it, every rank joins."""
    import torch
    import torch.distributed as dist
    from . import _lib
    lib = _lib.load()
    uid = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        _lib.check(lib.fpb_comm_unique_id(uid.ctypes.data))
    t = torch.from_numpy(uid)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    uid = t.cpu().numpy()
    _lib.check(lib.fpb_comm_init(op.h, uid.ctypes.data, world, rank), op.h)


def comm_kind(op) -> str:
    """How the handle sums its shards: 'none', 'nccl' (ncclAllReduce) or 'peer' (one kernel over
    NVLink peer memory, fused with the op's finalize step; csrc/fpb_peer.cuh)."""
    from . import _lib
    return ("none", "nccl", "peer")[int(_lib.load().fpb_comm_kind(op.h))]


def link_local(ops) -> None:
    """Shards held by operators of THIS process (one GPU, or GPUs with peer access) become ranks
    0..n-1; no NCCL.  Calls on the linked operators must run concurrently, one thread each
    (run_ranks): every rank's kernel waits for the others."""
    from . import _lib
    lib = _lib.load()
    arr = (ctypes.c_void_p * len(ops))(*[op.h for op in ops])
    _lib.check(lib.fpb_comm_link_local(arr, len(ops)), ops[0].h)


def run_ranks(ops, fn):
    """fn(op) on one thread per linked operator (ctypes calls release the GIL); results by rank."""
    import threading
    out, exc = [None] * len(ops), [None] * len(ops)

    def work(i):
        try:
            out[i] = fn(ops[i])
        except BaseException as e:  # noqa: BLE001
            exc[i] = e
    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(ops))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in exc:
        if e is not None:
            raise e
    return out


def peer_partition(count: int, world: int, grid: int):
    """Index arithmetic of the peer-memory shard sum (csrc/fpb_peer.cuh): rank r sums slice r =
    [r S, (r+1) S) with S = ceil(count / world); CTA b of a rank handles sub-slice
    [g S + b T, g S + (b+1) T) of every slice g, T = ceil(S / grid).  Returns
    (slices, subs): slices[r] = (lo, hi), subs[g][b] = (lo, hi) (empty ranges have lo >= hi)."""
    S = -(-count // world)
    T = -(-S // grid)
    slices = [(min(r * S, count), min((r + 1) * S, count)) for r in range(world)]
    subs = [[(g * S + b * T, min(g * S + (b + 1) * T, (g + 1) * S, count)) for b in range(grid)]
            for g in range(world)]
    return slices, subs


def two_shot_sum(parts):
    """Host model of the kernel's result: every slice is summed by one rank over the ranks' partial
    vectors in rank order and handed to all -- what every rank ends up holding."""
    world, count = len(parts), parts[0].shape[0]
    slices, _ = peer_partition(count, world, 1)
    out = np.empty(count)
    for lo, hi in slices:
        acc = parts[0][lo:hi].copy()
        for g in range(1, world):
            acc += parts[g][lo:hi]
        out[lo:hi] = acc
    return out
