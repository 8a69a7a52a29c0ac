#!/bin/bash
# usage: gpu_r2_mix.sh <tag> -- e2e legs (composition only), a small C4 run on one GPU, build-parameter sweep of cov_stream2_kernel
tag=${1:-x}
mkdir -p gpurun_out
MP2_TRACE=1 timeout 900 python bench.py --samples 0 --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; grep mp2_tnf_packed gpurun_out/bench_$tag.err | tail -4
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
e=d["e2e"]; print("E2E runs", round(e["value"],1), e["ms_per_step_min_median_max"], "mask", round(e["with_validity_mask_0.375_B_per_base"]["value"],1), "ascii", round(e["ascii_upload"]["value"],1))
PY
timeout 900 python bench.py --config c4 --c4-mags 300 --samples 3 --steps 2 --warmup 3 --no-cpu --no-cov-legs > gpurun_out/bench_${tag}_c4small.json 2> gpurun_out/bench_${tag}_c4small.err; echo "c4 small exit $?"; tail -2 gpurun_out/bench_${tag}_c4small.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_c4small.json")); c=d["coverage"]
print("C4 small", d["scaling"], round(d["value"],1), d["config"]["workload"], d["config"]["c4"], "cov %.3e" % c["value"], c["config"]["resident_event_sets"])
PY
run() { timeout 600 python bench.py --steps 3 --warmup 2 --samples 3 --no-cpu --no-cov-legs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['coverage']; print('$1', c['kernel_ms_per_sample'])"; }
run "base ST=64 NST=4"
for v in "-DMP2_C4_STAGE=128 -DMP2_C4_NST=4" "-DMP2_C4_STAGE=64 -DMP2_C4_NST=8" "-DMP2_C4_STAGE=128 -DMP2_C4_NST=2"; do
  make -C magpurify2_b200/csrc -B EXTRA="$v" mp2_cov.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1
  run "$v"
done
