"""world_size-2 gloo worker for test_host.py::test_two_rank_gloo_shard_sum."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_fixture  # noqa: E402
from flashpca_b200 import dist as fdist  # noqa: E402
from oracle import oracle as O  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
_, payload, n, p = load_fixture("data_chr1")
npb = (n + 3) // 4
j0, j1 = fdist.shard_range(p, world, rank)
assert fdist.shard_range(p, world, 0)[0] == 0 and fdist.shard_range(p, world, world - 1)[1] == p
local = O.COracle(payload[j0 * npb: j1 * npb], n, j1 - j0)
x = np.random.default_rng(0).standard_normal(n)
y = torch.from_numpy(local.perform_op(x, 0))
dist.all_reduce(y)
tr = torch.tensor([local.trace], dtype=torch.float64)
dist.all_reduce(tr)
full = O.COracle(payload, n, p)
yf = full.perform_op(x, 0)
ok = np.abs(y.numpy() - yf).max() <= 1e-12 * np.abs(yf).max() and abs(tr.item() - full.trace) <= 1e-12 * full.trace
msd = fdist.gather_meansd(local.meansd(), p, world, rank)
ok = ok and np.array_equal(msd, full.meansd())
# the library's own shard sum (csrc/fpb_peer.cuh) modelled on the host: every rank sees all partial
# vectors (P2P loads), slice r is summed in rank order by rank r and handed to everybody (P2P stores)
mine = torch.from_numpy(local.perform_op(x, 0))
parts = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(parts, mine)
slices, _ = fdist.peer_partition(n, world, 128)
lo, hi = slices[rank]
acc = parts[0][lo:hi].clone()
for g in range(1, world):
    acc += parts[g][lo:hi]
S = slices[0][1] - slices[0][0]                      # gloo gathers equal sizes: pad the last slice
pieces = [torch.empty(S, dtype=torch.float64) for _ in slices]
dist.all_gather(pieces, torch.cat([acc, torch.zeros(S - acc.numel(), dtype=torch.float64)]))
y2 = torch.cat([q[: s[1] - s[0]] for q, s in zip(pieces, slices)]).numpy()
ok = ok and np.array_equal(y2, fdist.two_shot_sum([q.numpy() for q in parts]))
ok = ok and np.abs(y2 - yf).max() <= 1e-12 * np.abs(yf).max()
ref = torch.from_numpy(y2.copy())
dist.broadcast(ref, src=0)
ok = ok and np.array_equal(ref.numpy(), y2)          # bit-identical on every rank
flag = torch.tensor([1 if ok else 0])
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0 and flag.item() == 1:
    print("SHARD_OK")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
