#!/bin/bash
# One gpurun call: fused-kernel tests, knob sweep, bench A/B, ncu captures, full GPU suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused" > gpurun_out/t_fused.log 2>&1
echo "fused tests rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/t_fused.log
timeout 600 python tools/fused_sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
echo "sweep rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/sweep.jsonl
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
FPB_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_two_kernel.json 2> gpurun_out/bench_two.err
cat gpurun_out/bench_fused.json
