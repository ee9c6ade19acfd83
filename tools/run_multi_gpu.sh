#!/bin/bash
# Multi-GPU session: bench at N GPUs (with the 1M x 500k side measurement at 8), the sharded NCCL test,
# and the all-reduce probe per NCCL algorithm.  Usage: tools/run_multi_gpu.sh <ngpus> [extra bench args]
N=${1:-8}; shift
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 "$@" > $OUT/bench_g${N}.json 2> $OUT/bench_g${N}.err
echo "bench rc=$?"; tail -c 400 $OUT/bench_g${N}.err
python - <<PY
import json
d=json.loads([l for l in open("$OUT/bench_g${N}.json") if l.startswith("{")][-1])
print("N", d["n_gpus"], "step_ms", d["step_ms"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["median_ms_per_step"])
print("roofline", d["roofline"]["frac"], d["roofline"]["halves_ms"], d["roofline"]["launch"]["launch_ms"])
s=d["solve"]; print("solve", {k:s[k] for k in ("seconds","iterate_seconds","eigenvector_download_seconds","nops","check_mse")}, s["eigenvalues_over_p"][:3])
c=d.get("config_1m_x_500k")
if c: print("cfg5", c["ms_per_step"], c["perform_op_frac_of_single_read_roofline"], c["stage_seconds"], {k:c["solve"][k] for k in ("seconds","nops","check_mse","nconv")}, c["solve"]["eigenvalues_over_p"][:3])
PY
timeout 200 python -m pytest tests/test_gpu_cli.py -q -k nccl 2>&1 | tail -3
for proto in LL128 Simple; do
  NCCL_PROTO=$proto timeout 120 $TR --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 --no-solve --no-cfg5 > $OUT/bench_g${N}_$proto.json 2>/dev/null
  python -c "
import json;d=json.loads([l for l in open('$OUT/bench_g${N}_$proto.json') if l.startswith('{')][-1]);print('NCCL_PROTO=$proto', d['step_ms'], 'e2e', d['e2e']['ms_per_step'])"
done
for s in "default::" "Ring:Simple:" "Ring:LL128:" "Tree:LL128:" "Tree:Simple:" "NVLS::" ":LL:"; do
  IFS=: read algo proto _ <<< "$s"
  [ "$algo" = default ] && algo=
  env ${algo:+NCCL_ALGO=$algo} ${proto:+NCCL_PROTO=$proto} timeout 60 $TR --master-port 29512 tools/allreduce_probe.py 2>/dev/null | tail -1 | sed "s/^/[$s] /" | tee -a $OUT/allreduce_probe_g${N}.txt
done
