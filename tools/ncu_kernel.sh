#!/bin/bash
# One `ncu --set full` capture of one kernel on an 8,192-particle bench run; run under gpurun.
# usage: bash tools/ncu_kernel.sh <tag> <kernel regex> [launch-skip]
TAG=$1; KRE=$2; SKIP=${3:-14}
ncu --set full --import-source on --clock-control none -k regex:${KRE} --launch-skip ${SKIP} --launch-count 1 \
    -o gpurun_out/${TAG} -f python bench.py --particles 8192 --steps 3 --warmup 3 --burnin 12 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}.ncu-rep
