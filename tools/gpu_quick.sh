#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?" | tee gpurun_out/summary.txt
tail -12 gpurun_out/t_gpu.log
cat > /tmp/exp.py <<'PY'
import ctypes, os, sys, json
sys.path.insert(0, os.getcwd())
import torch
from flashpca_b200 import _lib
from flashpca_b200.synth import SynthSpec
lib = _lib.load()
n, p = 500000, 100000
op = SynthSpec(n, p).create_operator(device=0)
k = 4
X = torch.randn(k * n, dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
V = torch.randn(k * p, dtype=torch.float64, device="cuda"); T = torch.empty(k * p, dtype=torch.float64, device="cuda")
stream = torch.cuda.ExternalStream(lib.fpb_stream(op.h))
def timed(fn, reps=5):
    fn(); _lib.check(lib.fpb_sync(op.h), op.h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(reps): fn()
        e1.record()
    _lib.check(lib.fpb_sync(op.h), op.h)
    return e0.elapsed_time(e1) / reps
res = {}
for pair in ("1", "0"):
    os.environ["FPB_PAIR"] = pair
PY
python - <<'PY'
print("see pair timing below")
PY
for PAIR in 1 0; do FPB_PAIR=$PAIR timeout 300 python - <<'PY'
import ctypes, os, sys, json
sys.path.insert(0, os.getcwd())
import torch
from flashpca_b200 import _lib
from flashpca_b200.synth import SynthSpec
lib = _lib.load()
n, p, k = 500000, 100000, 4
op = SynthSpec(n, p).create_operator(device=0)
X = torch.randn(k * n, dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
V = torch.randn(k * p, dtype=torch.float64, device="cuda"); T = torch.empty(k * p, dtype=torch.float64, device="cuda")
stream = torch.cuda.ExternalStream(lib.fpb_stream(op.h))
def timed(fn, reps=5):
    fn(); _lib.check(lib.fpb_sync(op.h), op.h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(reps): fn()
        e1.record()
    _lib.check(lib.fpb_sync(op.h), op.h)
    return e0.elapsed_time(e1) / reps
out = dict(pair=os.environ["FPB_PAIR"], k=k,
  perform_op_multi_ms=timed(lambda: _lib.check(lib.fpb_perform_op_multi_dev(op.h, X.data_ptr(), k, Y.data_ptr()), op.h)),
  crossprod_multi_ms=timed(lambda: _lib.check(lib.fpb_crossprod_multi_dev(op.h, X.data_ptr(), k, T.data_ptr()), op.h)),
  prod_multi_ms=timed(lambda: _lib.check(lib.fpb_prod_multi_dev(op.h, V.data_ptr(), k, Y.data_ptr()), op.h)))
print(json.dumps(out))
PY
done | tee gpurun_out/pair_timing.txt
