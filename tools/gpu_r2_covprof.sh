#!/bin/bash
# usage: gpu_r2_covprof.sh <tag> [pattern] -- coverage parity tests, a 3-sample C3 bench, one full ncu capture of the streaming kernel
tag=${1:-x}; pat=${2:-cov_stream2_kernel}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cov_gpu.py -m gpu -x -q 2>&1 | tail -3
ARGS="--steps 3 --warmup 2 --samples 3 --no-cpu --no-cov-legs"
timeout 600 python bench.py $ARGS > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -2 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json")); c=d["coverage"]
print("COV %.3e reads/s ms/step %.2f" % (c["value"], c["ms_per_step"]), c["kernel_ms_per_sample"])
PY
timeout 600 python bench.py $ARGS --max-span 150 > gpurun_out/bench_${tag}_v.json 2>> gpurun_out/bench_$tag.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 3 -c 1 -o gpurun_out/prof_$tag python bench.py $ARGS --max-span 150 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-200
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_v.json")); c=d["coverage"]
print("vouched COV %.3e reads/s ms/step %.2f" % (c["value"], c["ms_per_step"]), c["kernel_ms_per_sample"])
PY
