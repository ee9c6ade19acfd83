#!/bin/bash
# Full GPU suite + default bench + launch list + config-1 (10k x 100k) numbers.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?" | tee gpurun_out/summary.txt
tail -4 gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_default.json
timeout 300 python bench.py --steps 50 --warmup 5 --n 10000 --p 100000 --no-cpu-baseline > gpurun_out/bench_cfg1_10k.json 2> gpurun_out/bench_cfg1.err
cat gpurun_out/bench_cfg1_10k.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-solve --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
