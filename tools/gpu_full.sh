#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
python - <<PY
import json
r=json.load(open("gpurun_out/bench_ref.json")); print("REF", round(r["value"],3), "Gbp/s;", "%.3e reads/s" % r["coverage"]["value"])
d=json.load(open("gpurun_out/bench.json")); c=d["coverage"]
print("TNF", round(d["value"],1), "Gbp/s ms", round(d["ms_per_step"],3), "kernel frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], d["clocks"])
print("COV %.3e reads/s ms %.2f e2e %.3e" % (c["value"], c["ms_per_step"], c["e2e"]["value"]), c["roofline"]["kernel"], round(c["roofline"]["frac"],3))
print(d.get("cpu_baseline"), c.get("cpu_baseline"))
PY
