#!/usr/bin/env python
"""Per-op exchange cost of the SNP-sharded path: ncclAllReduce(sum, f64, N) on N = 500,000 and
1,000,000 doubles, for the NCCL_ALGO / NCCL_PROTO of the environment, device-timed with CUDA events (max over ranks).
Run: torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/allreduce_probe.py"""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world, "cases": []}
    name = "%s/%s" % (os.environ.get("NCCL_ALGO", "default"), os.environ.get("NCCL_PROTO", "default"))
    for n in (8, 500000, 1000000):
        x = torch.randn(n, dtype=torch.float64, device="cuda")
        for _ in range(5):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            dist.all_reduce(x)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50 * 1e3], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["cases"].append({"setting": name, "doubles": n, "us": round(t.item(), 2)})
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
