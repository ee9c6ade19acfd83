"""Peer-memory shard sum (csrc/fpb_peer.cuh): the SNP-sharded op of SURVEY section 8e with the
exchange done by the library's own kernel instead of ncclAllReduce.  Two or three shards of one
matrix live on ONE GPU here (fpb_comm_link_local), so the protocol -- flags, fixed-order sum,
fused finalize, chunked block form, CUDA-graph replay -- is exercised on a single-GPU box; the
multi-process form (CUDA IPC over fpb_comm_init) is tests/test_gpu_cli.py::test_nccl_sharded_two_gpus.

Ranks that share a process share a CUDA context: a device-synchronising call (cudaMalloc) issued
for one rank while another rank's kernel waits for it would never return, so every buffer is
allocated by a warm-up call before the shards are linked, and the device-pointer entry points
are enqueued for all ranks before any of them is waited for.  (The solver synchronises with the
host inside every step, so a solve over linked shards needs one process per rank: that is the
torchrun test.)"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
os.environ.setdefault("FPB_PEER_TIMEOUT_S", "5")   # a protocol failure ends the test, not the box


def _shards(spec, world, warm):
    from flashpca_b200 import dist as fdist
    ops = []
    for r in range(world):
        j0, j1 = fdist.shard_range(spec.p, world, r)
        ops.append(spec.create_operator(j0=j0, j1=j1))
    for op in ops:
        warm(op)
    fdist.link_local(ops)
    assert all(fdist.comm_kind(op) == "peer" for op in ops)
    return ops


@pytest.mark.parametrize("world", [2, 3])
def test_linked_shards_sum_like_the_whole_matrix(native_lib, world):
    import torch
    from flashpca_b200 import _lib
    from flashpca_b200.synth import SynthSpec
    lib = native_lib
    spec = SynthSpec(20011, 3000 * world + 7, seed=11, fst=0.05)
    n, k = spec.n, 11
    full = spec.create_operator()
    rng = np.random.default_rng(5)
    x = rng.standard_normal(n)
    m = np.asfortranarray(rng.standard_normal((n, k)))
    yf, Yf = full.perform_op(x), full.perform_op_mat(m)

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()   # column-major on the device

    xd, md = dev(x), dev(m)
    ys = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(world)]
    Ys = [torch.empty(k * n, dtype=torch.float64, device="cuda") for _ in range(world)]

    def warm(op):   # allocates the block-path scratch of the handle
        _lib.check(lib.fpb_perform_op_multi_dev(op.h, md.data_ptr(), k, Ys[0].data_ptr()), op.h)
        _lib.check(lib.fpb_sync(op.h), op.h)

    ops = _shards(spec, world, warm)
    torch.cuda.synchronize()
    # single vector: finalize + sum in one kernel; the same pointers recur, so the later
    # repetitions replay the op as a CUDA graph
    for rep in range(5):
        for y in ys:
            y.zero_()
        torch.cuda.synchronize()
        for op, y in zip(ops, ys):
            _lib.check(lib.fpb_perform_op_dev(op.h, xd.data_ptr(), y.data_ptr()), op.h)
        for op in ops:
            _lib.check(lib.fpb_sync(op.h), op.h)
        for y in ys:
            yh = y.cpu().numpy()
            assert np.abs(yh - yf).max() <= 1e-12 * np.abs(yf).max(), rep
            assert torch.equal(y, ys[0])                 # bit-identical on every rank
    # block form: N x k summed in chunks of 8 columns, k not a multiple of the chunk
    for op, Y in zip(ops, Ys):
        _lib.check(lib.fpb_perform_op_multi_dev(op.h, md.data_ptr(), k, Y.data_ptr()), op.h)
    for op in ops:
        _lib.check(lib.fpb_sync(op.h), op.h)
    for Y in Ys:
        Yh = Y.cpu().numpy().reshape(k, n).T
        assert np.abs(Yh - Yf).max() <= 1e-12 * np.abs(Yf).max()
        assert torch.equal(Y, Ys[0])
