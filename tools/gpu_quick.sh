#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?" | tee gpurun_out/summary.txt
tail -4 gpurun_out/t_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read())
print("ms/op", d['ms_per_step'], "e2e", d['e2e']['ms_per_step'], d['roofline']['launch_ms'], d['clocks'])
print("solve", d['solve'])
print("small", d['config_10k_x_100k'])
print("block", d['block_variant'])
print("cpu", d['cpu_baseline'])
PY
