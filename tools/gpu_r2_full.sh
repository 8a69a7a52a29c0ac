#!/bin/bash
# usage: gpu_r2_full.sh <tag> -- all GPU parity tests, smoke, the default bench line (C3) with all its legs
tag=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0
timeout 1500 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $? after ${SECONDS}s"; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json")); c=d["coverage"]
print("TNF", round(d["value"],1), "Gbp/s ms", round(d["ms_per_step"],3), d["ms_per_step_min_median_max"], "kernel frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"], d["clocks"])
e=d["e2e"]; print("E2E packed", round(e["value"],1), e["ms_per_step_min_median_max"], "ascii", round(e["ascii_upload"]["value"],1), "pack GB/s", round(e["host_pack"]["gb_per_s"],1), "threads", e["host_pack"]["threads"])
print("COV %.3e reads/s ms %.2f e2e %.3e" % (c["value"], c["ms_per_step"], c["e2e"]["value"]), c["roofline"]["kernel"], round(c["roofline"]["frac"],3), round(c["roofline"].get("moved_frac_of_peak",0),3), c["kernel_ms_per_sample"])
print("legs", json.dumps(c["legs"])[:1200])
print(d.get("cpu_baseline"), c.get("cpu_baseline"))
print(d["kernel_ms_per_step"], d["packed_2bit_path"])
PY
