#!/bin/bash
# Round-2 final artefacts on one GPU: GPU suite, smoke, default bench, reference arm, launch list of the bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/t_gpu_final.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/t_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err
echo "bench reference rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-solve --no-cpu-baseline --no-small --no-block > gpurun_out/ncu_bench_final.log 2>&1
echo "ncu launches rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final.json"))
s=d["solve"]
print("op", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "solve", s["seconds"], s["block_krylov"]["seconds"], "small", d["config_10k_x_100k"]["ms_per_step"], d["config_10k_x_100k"]["solve"]["seconds"], "blk20", d["block_variant"]["k20"]["ms_per_call"], "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
r=json.load(open("gpurun_out/bench_reference_final.json")); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
