#!/bin/bash
# usage: gpu_prof.sh <tag> <kernel-regex> [bench args...]
tag=$1; shift; rx=$1; shift
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu "$@" > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 1 --no-cpu "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-300
