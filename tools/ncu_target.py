#!/usr/bin/env python
"""Short target for ncu: two perform_op and one 8-column block op at 500,000 x 100,000.
  ncu --set full --clock-control none --import-source on -k regex:'k_umma|k_imma_gemv_tma' -c 8 \
      -o gpurun_out/r02_kernels python tools/ncu_target.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("FPB_GRAPH", "0")
from flashpca_b200.synth import SynthSpec  # noqa: E402

n, p = 500000, 100000
op = SynthSpec(n, p).create_operator()
rng = np.random.default_rng(0)
x = rng.standard_normal(n)
for _ in range(2):
    op.perform_op(x)
m = rng.standard_normal((n, 8))
op.perform_op_mat(m)
op.close()
print("done")
