#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_cov_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -15
MP2_COV_PATH=general python -m pytest tests/test_cov_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cov_tile.json 2> gpurun_out/bench_cov_tile.err; echo exit $?; tail -3 gpurun_out/bench_cov_tile.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_cov_tile.json")); c=d["coverage"]
print("tnf", round(d["value"],1), round(d["roofline"]["frac"],3))
print("cov reads/s %.3e ms/step %.2f" % (c["value"], c["ms_per_step"]))
for k in c["kernels"]: print(k["kernel"], round(k["kernel_ms"],3), "ms", round(k["achieved"],1), "GB/s")
print(c["e2e"]["value"], c.get("cpu_baseline"))
PY
