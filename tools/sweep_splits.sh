mkdir -p gpurun_out
run() { (env "$@" timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-solve) 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$*', round(j['ms_per_step'],3), {k:round(v,3) for k,v in j['roofline']['launch_ms'].items()})"; }
for s in 2 3 4; do run FPB_DEBUG_SPLITS1=$s FPB_DEBUG_SPLITS2=8; done
for t in 12 16 24 48; do run FPB_DEBUG_SPLITS1=3 FPB_DEBUG_SPLITS2=$t; done
