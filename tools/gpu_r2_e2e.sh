#!/bin/bash
# usage: gpu_r2_e2e.sh <tag> -- packed-upload parity tests + composition-only bench (e2e legs)
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_tnf_gpu.py -m gpu -x -q 2>&1 | tail -4
MP2_TRACE=1 timeout 900 python bench.py --samples 0 --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
e=d["e2e"]; print("E2E runs", round(e["value"],1), e["ms_per_step_min_median_max"], "mask", round(e["with_validity_mask_0.375_B_per_base"]["value"],1), "ascii", round(e["ascii_upload"]["value"],1), "pack GB/s", round(e["host_pack"]["gb_per_s"],1), e["invalid_runs"], "h2d", e["h2d_bytes_per_step"])
print("TNF", round(d["value"],1), d["ms_per_step_min_median_max"], round(d["roofline"]["frac"],3))
PY
