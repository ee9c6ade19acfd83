#!/bin/bash
# multi-GPU check: NCCL tests + strong-scaling bench at the visible GPU count
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus=$N" | tee gpurun_out/multi_summary.txt
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -m gpu -k "nccl or shard or multi" > gpurun_out/t_multi.log 2>&1
echo "multi tests rc=$?" | tee -a gpurun_out/multi_summary.txt
tail -3 gpurun_out/t_multi.log
for P in unset 0; do
  if [ "$P" = "0" ]; then export FPB_PERSIST=0; else unset FPB_PERSIST; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g${N}_persist_$P.json 2> gpurun_out/bench_g${N}_$P.err
  echo "bench gpus=$N persist=$P rc=$?" | tee -a gpurun_out/multi_summary.txt
  python - <<PY
import json
for ln in open('gpurun_out/bench_g${N}_persist_$P.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['n_gpus'], d['ms_per_step'], d['roofline']['launch_ms'], d['e2e']['ms_per_step'], d['solve']['seconds'], d['solve']['nops'])
PY
done
