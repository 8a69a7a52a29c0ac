#!/bin/bash
# Round-1 evidence: (1) plain bench, (2) launch list of our kernels with device times, (3) one
# --set full capture of the two dominant kernels at the FULL bench size (C3).
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 3 --no-cpu"
RX='regex:tnf_|cov_|clr_|wmedian|logratio|select_samples'
python bench.py $ARGS > gpurun_out/plain_all.json 2> gpurun_out/plain_all.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" -c 700 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
python bench.py $ARGS > gpurun_out/plain_all2.json 2> gpurun_out/plain_all2.err &&
ncu --set full --clock-control none --import-source on -k regex:'^tnf_kernel|cov_stream_kernel' -s 4 -c 2 -o gpurun_out/prof_full python bench.py $ARGS > gpurun_out/ncu_full_all.log 2>&1
tail -2 gpurun_out/ncu_full_all.log | cut -c1-160
