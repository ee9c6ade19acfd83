#!/usr/bin/env python
"""ncu target for k_allreduce_peer: one 8-GPU shard (500,000 x 12,500) linked as a communicator of
ONE rank (no peers to wait for: ncu serialises kernels, so a real exchange cannot be profiled) --
shows the local side of the fused kernel (finalize from the split partials + copy-out).
  ncu --set full --clock-control none -k regex:'k_allreduce_peer|k_finalize_prod' -c 6 python tools/ncu_peer_target.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("FPB_GRAPH", "0")
from flashpca_b200 import dist as fdist  # noqa: E402
from flashpca_b200.synth import SynthSpec  # noqa: E402

n, p = 500000, 100000
x = np.random.default_rng(0).standard_normal(n)
plain = SynthSpec(n, p).create_operator(j0=0, j1=12500)
for _ in range(2):
    y0 = plain.perform_op(x)            # k_finalize_prod
plain.close()
op = SynthSpec(n, p).create_operator(j0=0, j1=12500)
fdist.link_local([op])
for _ in range(3):
    y1 = op.perform_op(x)               # k_allreduce_peer<kPeerFinalize> (+ kPeerGather for the upload)
print("same result:", bool(np.array_equal(y0, y1)))
