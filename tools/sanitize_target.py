#!/usr/bin/env python
"""Small workload that touches every kernel family once, for compute-sanitizer:
  compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_target.py
(fixture data_chr1: 957 x 1129; ragged tails in every tile dimension)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flashpca_b200 import SVDWide, SVDWideOnline  # noqa: E402
from oracle import oracle as O  # noqa: E402

stem = os.path.join(ROOT, "tests", "golden", "data_chr1", "data_chr1")
n = O.count_lines(stem + ".fam")
payload, _, p = O.read_bed_payload(stem + ".bed", n)
rng = np.random.default_rng(0)
x = rng.standard_normal(n)
for env in ({}, {"FPB_PATH": "generic"}, {"FPB_GEMV": "ldg"}, {"FPB_FUSED": "1"}, {"FPB_PERSIST": "0"},
            {"FPB_GATHER_SMS": "2"}):
    for k in ("FPB_PATH", "FPB_GEMV", "FPB_FUSED", "FPB_PERSIST", "FPB_GATHER_SMS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    op = SVDWideOnline(payload=payload, n=n, nsnps=p)
    y = op.perform_op(x)
    t = op.crossprod(x)
    z = op.prod(t)
    print(env, "op vs prod(crossprod): %.2e" % (np.abs(y - z).max() / np.abs(y).max()), flush=True)
    if not env:
        m = rng.standard_normal((n, 11))
        Y = op.perform_op_mat(m)          # tcgen05 (8) + pair (2) + single (1)
        print("block vs single: %.2e" % (np.abs(Y[:, 10] - op.perform_op(m[:, 10])).max() / np.abs(Y).max()))
        r = op.pca(5, 11, 100, 1e-6)
        rb = op.pca_block(5, 1e-6)
        print("pca nconv", r["nconv"], "block nconv", rb["nconv"], flush=True)
    op.close()
d = SVDWide(rng.integers(0, 3, size=(300, 200)).astype(np.float64), 3)
print("dense", float(np.abs(d.perform_op(rng.standard_normal(300))).max()))
