#!/bin/bash
# ncu --set full capture of the fused kernel (second launch), summarised on the box.
mkdir -p gpurun_out
W=${FPB_FUSED_WINDOW:-4}
FPB_FUSED_WINDOW=$W timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_op -s 1 -c 1 \
  -o gpurun_out/fused_w$W -f python tools/ncu_fused.py > gpurun_out/ncu_fused.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fused.log
python tools/ncu_summary.py gpurun_out/fused_w$W.ncu-rep gpurun_out/ncu_fused_w$W.txt
ncu -i gpurun_out/fused_w$W.ncu-rep --page details > gpurun_out/ncu_fused_w${W}_details.txt 2>&1
grep -E "time_duration|dram__bytes_read.sum |dram__bytes_write.sum |hit_rate|tensor_subpipe_imma|issue_active|wavefronts_mem_shared.sum.pct|lts__t_bytes|throughput" gpurun_out/ncu_fused_w$W.txt | head -40
