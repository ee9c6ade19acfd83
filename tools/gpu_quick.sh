#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?" | tee gpurun_out/summary.txt
tail -5 gpurun_out/t_gpu.log
cat > /tmp/exp.py <<'PY'
import ctypes, os, sys, json
sys.path.insert(0, os.getcwd())
import torch
from flashpca_b200 import _lib
from flashpca_b200.synth import SynthSpec
lib = _lib.load()
n, p = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn(n, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
ms = ctypes.c_float(); kms = (ctypes.c_float * 4)()
ref = None
for env in [dict(FPB_P2WIDE="0"), dict()] + [dict(FPB_DEBUG_SPLITS2=str(b)) for b in (4, 6, 12, 16)]:
    for k in ("FPB_P2WIDE", "FPB_DEBUG_SPLITS2"): os.environ.pop(k, None)
    os.environ.update(env)
    op = SynthSpec(n, p).create_operator(device=0)
    _lib.check(lib.fpb_time_perform_op(op.h, x.data_ptr(), y.data_ptr(), 3, ctypes.byref(ms), None), op.h)
    _lib.check(lib.fpb_time_perform_op(op.h, x.data_ptr(), y.data_ptr(), 20, ctypes.byref(ms), kms), op.h)
    torch.cuda.synchronize()
    if ref is None: ref = y.clone()
    print(json.dumps(dict(n=n, p=p, env=env, ms=ms.value, k1=kms[2], k2=kms[3], relerr=float((y-ref).abs().max()/ref.abs().max()))), flush=True)
    op.close()
PY
timeout 600 python /tmp/exp.py 500000 100000 | tee gpurun_out/wide_sweep.txt
