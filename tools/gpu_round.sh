#!/bin/bash
# One gpurun call: fused-kernel tests, protocol latency, knob sweep, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused" > gpurun_out/t_fused.log 2>&1
echo "fused tests rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/t_fused.log
for w in 2 4; do FPB_FUSED_WINDOW=$w timeout 300 python tools/fused_latency.py > gpurun_out/latency_w$w.txt 2>&1; cat gpurun_out/latency_w$w.txt; done
timeout 600 python tools/fused_sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
echo "sweep rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/sweep.jsonl
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_fused.json
