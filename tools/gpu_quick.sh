#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1
echo "gpu tests rc=$?" | tee gpurun_out/summary.txt
tail -25 gpurun_out/t_gpu.log
