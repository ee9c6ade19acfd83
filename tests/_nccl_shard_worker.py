"""torchrun worker (N >= 2 GPUs): SNP-sharded perform_op / fpb_pca with the
library's NCCL all-reduce against the un-sharded single-GPU operator."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flashpca_b200 import dist as fdist  # noqa: E402
from flashpca_b200.synth import SynthSpec  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
spec = SynthSpec(20011, 30000, seed=11, fst=0.05)
j0, j1 = fdist.shard_range(spec.p, world, rank)
op = spec.create_operator(device=local, j0=j0, j1=j1)
fdist.attach_nccl(op, world, rank)
kind = fdist.comm_kind(op)
if os.environ.get("FPB_PEER", "1") != "0":
    assert kind == "peer", "peer-memory exchange could not be set up (%s)" % kind
x = np.random.default_rng(3).standard_normal(spec.n)
y = op.perform_op(x)                      # summed over the shards inside the library
m = np.asfortranarray(np.random.default_rng(4).standard_normal((spec.n, 11)))
Y = op.perform_op_mat(m)                  # block form: 8 + 3 columns
res = op.pca(10, 21, 500, 1e-8)
tr = torch.tensor([op.trace], dtype=torch.float64, device="cuda")
dist.all_reduce(tr)
ok = True
if rank == 0:
    full = spec.create_operator(device=local)
    yf = full.perform_op(x)
    rf = full.pca(10, 21, 500, 1e-8)
    e1 = np.abs(y - yf).max() / np.abs(yf).max()
    Yf = full.perform_op_mat(m)
    e1 = max(e1, np.abs(Y - Yf).max() / np.abs(Yf).max())
    e2 = np.abs(res["values"] / rf["values"] - 1).max()
    e3 = abs(tr.item() / full.trace - 1)
    ok = e1 < 1e-12 and e2 < 1e-9 and e3 < 1e-12 and res["nconv"] == 10
    print("shard sum: %s" % kind)
    print("sharded vs single: op %.2e eig %.2e trace %.2e nops %d/%d" % (e1, e2, e3, res["nops"], rf["nops"]))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
# every rank ran the Lanczos driver redundantly and must agree bit for bit
vals = torch.from_numpy(res["values"]).cuda()
ref = vals.clone()
dist.broadcast(ref, src=0)
same = torch.tensor([1 if torch.equal(vals, ref) else 0], device="cuda")
dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0 and flag.item() == 1 and same.item() == 1:
    print("NCCL_SHARD_OK")
dist.destroy_process_group()
sys.exit(0 if (flag.item() == 1 and same.item() == 1) else 1)
