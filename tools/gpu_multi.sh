#!/bin/bash
# multi-GPU check: NCCL tests + strong-scaling bench at the visible GPU count
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus=$N" | tee gpurun_out/multi_summary.txt
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -m gpu -k "nccl or shard or multi" > gpurun_out/t_multi.log 2>&1
echo "multi tests rc=$?" | tee -a gpurun_out/multi_summary.txt
tail -3 gpurun_out/t_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_g${N}.json 2> gpurun_out/bench_g${N}.err
echo "bench gpus=$N rc=$?" | tee -a gpurun_out/multi_summary.txt
python - <<PY
import json
for ln in open('gpurun_out/bench_g${N}.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['n_gpus'], d['ms_per_step'], d['roofline']['launch_ms'], d['e2e']['ms_per_step'], d['solve']['seconds'], d['solve']['nops'], d['clocks'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | tail -1 | cut -c1-300
