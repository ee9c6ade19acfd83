#!/bin/bash
# ncu --set full capture of the fused kernel (second launch), summarised on the box.
mkdir -p gpurun_out
W=${FPB_FUSED_WINDOW:-4}
export FPB_FUSED=1
FPB_FUSED_WINDOW=$W timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_op -s 1 -c 1 \
  -o gpurun_out/fused_w$W -f python tools/ncu_fused.py > gpurun_out/ncu_fused.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_fused.log
python tools/ncu_summary.py gpurun_out/fused_w$W.ncu-rep gpurun_out/ncu_fused_w$W.txt
ncu -i gpurun_out/fused_w$W.ncu-rep --page details > gpurun_out/ncu_fused_w${W}_details.txt 2>&1
grep -E "time_duration|dram__bytes_read.sum |dram__bytes_write.sum |hit_rate|tensor_subpipe_imma|issue_active|wavefronts_mem_shared.sum.pct|lts__t_bytes|throughput" gpurun_out/ncu_fused_w$W.txt | head -40
ncu -i gpurun_out/fused_w$W.ncu-rep --page source --csv > gpurun_out/fused_source.csv 2>/dev/null
python3 - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/fused_source.csv')))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[ix[k]].replace(',',''))
    except: return 0.0
tot=sum(f(r,'# Samples') for r in data)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print("total samples",tot)
agg={s:sum(f(r,s) for r in data) for s in stalls}
print(sorted(((v,k) for k,v in agg.items()),reverse=True)[:8])
for r in sorted(data,key=lambda r:-f(r,'# Samples'))[:30]:
    st=sorted(((f(r,s),s) for s in stalls),reverse=True)[:2]
    print("%6d %5.1f%% %-60s %s"%(data.index(r),100*f(r,'# Samples')/tot,r[ix['Source']][:60]," ".join("%s=%d"%(s[6:],v) for v,s in st)))
PY
