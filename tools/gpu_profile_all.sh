#!/bin/bash
# ncu launch list + one full capture per top kernel, on a reduced C3 shape (200 MAGs, 3 samples)
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 1 --no-cpu --mags 200 --samples 3"
python bench.py $ARGS > gpurun_out/plain_all.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
python bench.py $ARGS > gpurun_out/plain_all2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'tnf_kernel|cov_depth_kernel|cov_scatter_kernel|cov_tile_kernel' -s 3 -c 3 -o gpurun_out/prof_all python bench.py $ARGS > gpurun_out/ncu_full_all.log 2>&1
tail -3 gpurun_out/ncu_full_all.log | cut -c1-200
