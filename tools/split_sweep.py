#!/usr/bin/env python
"""Contraction-kernel times of one 8-GPU shard (500,000 x 12,500) for different split counts
(FPB_DEBUG_SPLITS1 = column splits of the first half, FPB_DEBUG_SPLITS2 = row splits of the second,
FPB_PERSIST = persistent forms).  One JSON line per setting."""
import ctypes
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from flashpca_b200 import _lib
    from flashpca_b200.synth import SynthSpec
    n, p = int(sys.argv[2]), int(sys.argv[3])
    lib = _lib.load()
    op = SynthSpec(n, 100000).create_operator(j0=0, j1=p)
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    ms = ctypes.c_float()
    kms = (ctypes.c_float * 4)()
    _lib.check(lib.fpb_time_perform_op(op.h, x.data_ptr(), y.data_ptr(), 5, ctypes.byref(ms), None), op.h)
    best = None
    for _ in range(5):
        _lib.check(lib.fpb_time_perform_op(op.h, x.data_ptr(), y.data_ptr(), 20, ctypes.byref(ms), kms), op.h)
        cur = (ms.value, kms[2], kms[3])
        best = cur if best is None or cur[0] < best[0] else best
    print(json.dumps({"s1": os.environ.get("FPB_DEBUG_SPLITS1"), "s2": os.environ.get("FPB_DEBUG_SPLITS2"),
                      "persist": os.environ.get("FPB_PERSIST"), "op_ms": round(best[0], 4),
                      "xtx_ms": round(best[1], 4), "xt_ms": round(best[2], 4)}))
    sys.exit(0)

n, p = 500000, 12500
settings = [({}, "default")]
for s1 in (8, 15, 30, 120):
    settings.append(({"FPB_DEBUG_SPLITS1": str(s1), "FPB_PERSIST": "0"}, "s1=%d" % s1))
for s2 in (1, 2, 4, 8):
    settings.append(({"FPB_DEBUG_SPLITS2": str(s2), "FPB_PERSIST": "0"}, "s2=%d" % s2))
settings.append(({"FPB_PERSIST": "1"}, "persist"))
settings.append(({"FPB_PERSIST": "0"}, "one-shot"))
for env, name in settings:
    e = dict(os.environ, FPB_GRAPH="0", **env)
    out = subprocess.run([sys.executable, __file__, "child", str(n), str(p)], env=e, capture_output=True, text=True)
    print(name, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
