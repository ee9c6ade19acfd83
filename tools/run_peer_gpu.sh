#!/bin/bash
# Peer-memory shard sum vs ncclAllReduce at N GPUs: the sharded tests (both kinds), the bench with
# the peer kernel (full line) and with FPB_PEER=0 (op only).  Usage: tools/run_peer_gpu.sh <ngpus>
N=${1:-2}; shift
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_cli.py -q -k nccl 2>&1 | tail -15
export FPB_PEER_TIMEOUT_S=20
timeout 300 $TR --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 "$@" > $OUT/bench_peer_g${N}.json 2> $OUT/bench_peer_g${N}.err
echo "bench peer rc=$?"; tail -c 600 $OUT/bench_peer_g${N}.err
FPB_PEER=0 timeout 200 $TR --master-port 29513 bench.py --gpus $N --steps 50 --warmup 5 --no-solve --no-cfg5 > $OUT/bench_nccl_g${N}.json 2> $OUT/bench_nccl_g${N}.err
echo "bench nccl rc=$?"
python - <<PY
import json
for tag in ("peer", "nccl"):
    try:
        d=json.loads([l for l in open("$OUT/bench_%s_g${N}.json" % tag) if l.startswith("{")][-1])
    except Exception as e:
        print(tag, "no line", e); continue
    print(tag, "N", d["n_gpus"], d["config"]["sharding"][:60], "step_ms", d["step_ms"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["median_ms_per_step"])
    print("  roofline", d["roofline"]["frac"], d["roofline"]["halves_ms"], d["roofline"]["launch"]["launch_ms"])
    s=d.get("solve")
    if s: print("  solve", {k:s[k] for k in ("seconds","iterate_seconds","eigenvector_download_seconds","nops","check_mse")}, s["eigenvalues_over_p"][:3])
    c=d.get("config_1m_x_500k")
    if c: print("  cfg5", c["ms_per_step"], c["perform_op_frac_of_single_read_roofline"], {k:c["solve"][k] for k in ("seconds","nops","check_mse","nconv")})
PY
