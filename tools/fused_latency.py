"""Per-slab time line of the fused kernel's cross-CTA protocol (FPB_FUSED_DEBUG=1).
Events: 0 producer-1 starts the slab's loads, 1 first-half warp 0 stored its partials,
3 reducer saw its row complete, 4 reducer done, 5 builder has every a_j, 6 a-slices built,
7 second-half warp 0 finished the slab."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FPB_FUSED_DEBUG"] = "1"
os.environ.setdefault("FPB_FUSED", "1")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from flashpca_b200 import _lib  # noqa: E402
from flashpca_b200.synth import SynthSpec  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
lib = _lib.load()
op = SynthSpec(n, p).create_operator(device=0)
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
ms = ctypes.c_float()
for _ in range(3):
    _lib.check(lib.fpb_time_perform_op(op.h, x.data_ptr(), y.data_ptr(), 1, ctypes.byref(ms), None),
               op.h)
buf = np.zeros((2, 256, 8), dtype=np.uint64)
_lib.check(lib.fpb_fused_debug(op.h, buf.ctypes.data, buf.size), op.h)
print("window", os.environ.get("FPB_FUSED_WINDOW", "default"), "ms/op", ms.value)
for c, name in enumerate(("cta0", "ctaLast")):
    t = buf[c].astype(np.int64)
    t0 = t[0, 0]
    sl = slice(16, 200)
    print(name, "slab period (us): ev0 %.2f ev1 %.2f ev7 %.2f" % tuple(
        np.diff(t[sl, e]).mean() / 1e3 for e in (0, 1, 7)))
    for e in (1, 3, 4, 5, 6, 7):
        d = (t[sl, e] - t[sl, 0]) / 1e3
        print("  ev%d - ev0: mean %.2f us  min %.2f  max %.2f" % (e, d.mean(), d.min(), d.max()))
    print("  first slabs (us since start):")
    for s in range(6):
        print("   ", s, " ".join("%8.2f" % ((t[s, e] - t0) / 1e3) for e in (0, 1, 3, 4, 5, 6, 7)))
