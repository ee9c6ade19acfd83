"""Driver for `ncu -k regex:k_fused_op`: stage the bench matrix, run a few perform_ops."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from flashpca_b200 import _lib  # noqa: E402
from flashpca_b200.synth import SynthSpec  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = _lib.load()
op = SynthSpec(n, p).create_operator(device=0)
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
for _ in range(reps):
    _lib.check(lib.fpb_perform_op_dev(op.h, x.data_ptr(), y.data_ptr()), op.h)
_lib.check(lib.fpb_sync(op.h), op.h)
print("ok", float(y.norm()))
