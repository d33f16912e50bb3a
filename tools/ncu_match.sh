#!/bin/bash
# One `ncu --set full` capture of match_kernel (source view) on an 8,192-particle launch; run under gpurun.
# usage: bash tools/ncu_match.sh <tag> [kernel regex]
TAG=${1:-match}
KRE=${2:-match_kernel}
python bench.py --particles 8192 --steps 3 --warmup 3 --burnin 12 --no-cpu-baseline > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --set full --import-source on --clock-control none -k regex:${KRE} --launch-skip 14 --launch-count 1 \
    -o gpurun_out/${TAG} -f python bench.py --particles 8192 --steps 3 --warmup 3 --burnin 12 --no-cpu-baseline > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}.ncu-rep
