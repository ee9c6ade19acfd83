"""Timing sweep of the fused perform_op kernel's knobs (window, L2 hints) on the
bench matrix.  Usage: python tools/fused_sweep.py [n p]"""
import ctypes
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from flashpca_b200 import _lib  # noqa: E402
from flashpca_b200.synth import SynthSpec  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
lib = _lib.load()
spec = SynthSpec(n, p)
x = torch.randn(n, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
ms = ctypes.c_float()
kms = (ctypes.c_float * 4)()
ref = None
configs = [dict(FPB_FUSED="0")]
for w, pf in itertools.product((3, 4, 5, 6, 8), (0, 1)):
    configs.append(dict(FPB_FUSED="1", FPB_FUSED_WINDOW=str(w), FPB_FUSED_PREFETCH=str(pf)))
for cfg in configs:
    for k in ("FPB_FUSED", "FPB_FUSED_WINDOW", "FPB_FUSED_POL1", "FPB_FUSED_POL2", "FPB_FUSED_PREFETCH"):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    op = spec.create_operator(device=0)
    try:
        _lib.check(lib.fpb_time_perform_op(op.h, x.data_ptr(), y.data_ptr(), 3, ctypes.byref(ms),
                                           None), op.h)
        _lib.check(lib.fpb_time_perform_op(op.h, x.data_ptr(), y.data_ptr(), 10, ctypes.byref(ms),
                                           kms), op.h)
        torch.cuda.synchronize()
        if ref is None:
            ref = y.clone()
            err = 0.0
        else:
            err = float((y - ref).abs().max() / ref.abs().max())
        print(json.dumps(dict(cfg=cfg, ms_per_op=ms.value, kernel_ms=[kms[i] for i in range(4)],
                              relerr_vs_two_kernel=err)), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps(dict(cfg=cfg, error=str(e))), flush=True)
    op.close()
