#!/bin/bash
# usage: gpu_r2_cov.sh <tag> -- coverage iteration of round 2: parity tests, C3 bench with the new streaming kernel
# (legs: order measured / vouched, indel variant) and the round-1 kernel for comparison.  Everything under timeout.
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cov_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -3 gpurun_out/bench_$tag.err
MP2_COV_STREAM=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-cov-legs --max-span 150 > gpurun_out/bench_${tag}_old.json 2> gpurun_out/bench_${tag}_old.err; echo "old exit $?"
python - <<PY
import json
for f in ("gpurun_out/bench_$tag.json", "gpurun_out/bench_${tag}_old.json"):
    try:
        d=json.load(open(f)); c=d["coverage"]
        print(f, "COV %.3e reads/s ms/step %.2f" % (c["value"], c["ms_per_step"]), c["roofline"]["kernel"], "kernel_ms %.3f" % c["roofline"]["kernel_ms"], "check_ms", round(c.get("order_check_ms_per_sample",0),3), c.get("cpu_baseline",{}).get("parity_with_gpu"))
        print("   legs", json.dumps(c.get("legs"))[:900])
        print("   TNF kernel_ms", d["roofline"]["kernel_ms"], "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1))
    except Exception as e: print("fail", f, e)
PY
