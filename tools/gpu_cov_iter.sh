#!/bin/bash
# usage: gpu_cov_iter.sh <tag>  -- coverage parity tests, C3 bench (with the CPU parity leg), one full ncu capture
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests/test_cov_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -2 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json')); c=d['coverage']
print("COV %.3e reads/s ms/step %.2f kernel_ms %.3f" % (c["value"], c["ms_per_step"], c["roofline"]["kernel_ms"]), c.get("cpu_baseline",{}).get("parity_with_gpu"))
PY
if [ -z "$NO_NCU" ]; then
ncu --set full --clock-control none --import-source on -k regex:'cov_stream_kernel' -s 4 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-200
fi
