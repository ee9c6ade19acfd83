#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/exp.py <<'PY'
import ctypes, os, sys, json
sys.path.insert(0, os.getcwd())
import torch
from flashpca_b200 import _lib
from flashpca_b200.synth import SynthSpec
lib = _lib.load()
op = SynthSpec(500000, 100000).create_operator(device=0)
x = torch.randn(500000, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
ms = ctypes.c_float(); kms = (ctypes.c_float * 4)()
_lib.check(lib.fpb_time_perform_op(op.h, x.data_ptr(), y.data_ptr(), 3, ctypes.byref(ms), None), op.h)
_lib.check(lib.fpb_time_perform_op(op.h, x.data_ptr(), y.data_ptr(), 10, ctypes.byref(ms), kms), op.h)
print(json.dumps(dict(env={k: v for k, v in os.environ.items() if k.startswith("FPB_")}, ms=ms.value, fused_ms=kms[2])))
PY
for mode in 0 1; do for w in 4 8; do FPB_FUSED=1 FPB_FUSED_DBGMODE=$mode FPB_FUSED_WINDOW=$w timeout 300 python /tmp/exp.py 2>&1 | tail -1; done; done | tee gpurun_out/exp.txt
