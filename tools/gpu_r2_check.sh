#!/bin/bash
# usage: gpu_r2_check.sh <tag> -- coverage tests, the default bench twice (run-to-run spread), C1/C2 from files
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cov_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
timeout 1500 python bench.py --no-cpu > gpurun_out/bench_${tag}_$i.json 2> gpurun_out/bench_${tag}_$i.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_$i.json")); c=d["coverage"]
print("TNF", round(d["value"],1), d["ms_per_step_min_median_max"], "e2e", round(d["e2e"]["value"],1), "COV %.3e ms %.2f second pass %.2f" % (c["value"], c["ms_per_step"], c["ms_per_step_one_stream"]), c["kernel_ms_per_sample"])
PY
done
timeout 600 python tools/run_c1_c2.py /tmp/mp2_c1c2 > gpurun_out/c1c2_$tag.json 2> gpurun_out/c1c2_$tag.err; tail -c 700 gpurun_out/c1c2_$tag.json; echo
