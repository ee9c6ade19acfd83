#!/bin/bash
# Final round artefacts: GPU suite, default bench, launch list, ncu --set full of the hot kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?" | tee gpurun_out/summary.txt
tail -4 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_default.json | cut -c1-3000
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-solve --no-cpu-baseline --no-small > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_imma_gemv_tma$|k_imma_gemv_tma_t$|k_imma_gemv_tma_2v|k_imma_gemv_tma_t_2v" -c 4 -o gpurun_out/final_kernels -f python - > gpurun_out/ncu_final.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from flashpca_b200 import _lib
from flashpca_b200.synth import SynthSpec
lib = _lib.load()
n, p = 500000, 100000
op = SynthSpec(n, p).create_operator(device=0)
x = torch.randn(2 * n, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
_lib.check(lib.fpb_perform_op_dev(op.h, x.data_ptr(), y.data_ptr()), op.h)
_lib.check(lib.fpb_perform_op_multi_dev(op.h, x.data_ptr(), 2, y.data_ptr()), op.h)
_lib.check(lib.fpb_sync(op.h), op.h)
PY
echo "ncu full rc=$?" | tee -a gpurun_out/summary.txt
python tools/ncu_summary.py gpurun_out/final_kernels.ncu-rep gpurun_out/ncu_final_kernels.txt
grep -E "^==|time_duration.sum |dram__bytes_read.sum |imma.avg|issue_active.avg|wavefronts_mem_shared.sum.pct" gpurun_out/ncu_final_kernels.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
