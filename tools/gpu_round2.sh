#!/bin/bash
# Round-2 evidence in one call: GPU parity tests, smoke, both bench arms (timed), C5, C1/C2 from files, launch list,
# one full ncu capture of the two dominant kernels at the full bench size (C3).  Outputs under gpurun_out/.
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_$tag.txt 2>&1
nproc >> gpurun_out/gpu_$tag.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/gpu_$tag.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log; tail -4 gpurun_out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0
timeout 900 python bench.py --impl reference > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; echo "ref exit $? after ${SECONDS}s"
SECONDS=0
timeout 1500 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $? after ${SECONDS}s"; tail -3 gpurun_out/bench_$tag.err
timeout 600 python bench.py --steps 3 --warmup 3 --long-contigs --mags 400 --samples 0 --no-cpu > gpurun_out/bench_${tag}_c5.json 2>> gpurun_out/bench_$tag.err
timeout 600 python tools/run_c1_c2.py /tmp/mp2_c1c2 > gpurun_out/c1c2_$tag.json 2> gpurun_out/c1c2_$tag.err; tail -c 900 gpurun_out/c1c2_$tag.json; echo
python - <<PY
import json
r=json.load(open("gpurun_out/bench_${tag}_ref.json")); print("REF", round(r["value"],3), "Gbp/s;", "%.3e reads/s" % r["coverage"]["value"])
d=json.load(open("gpurun_out/bench_$tag.json")); c=d["coverage"]
print("TNF", round(d["value"],1), "Gbp/s ms", round(d["ms_per_step"],3), d["ms_per_step_min_median_max"], "kernel frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "ascii", round(d["e2e"]["ascii_upload"]["value"],1), "launches", d["gpu_launches"], d["clocks"])
print("COV %.3e reads/s ms %.2f e2e %.3e" % (c["value"], c["ms_per_step"], c["e2e"]["value"]), c["roofline"]["kernel"], round(c["roofline"]["frac"],3), round(c["roofline"].get("moved_frac_of_peak",0),3), c["kernel_ms_per_sample"])
print("legs", json.dumps(c["legs"])[:1500])
print(d.get("cpu_baseline"), c.get("cpu_baseline"))
d=json.load(open("gpurun_out/bench_${tag}_c5.json")); print("C5", round(d["value"],1), "kernel frac", round(d["roofline"]["frac"],3))
PY
ARGS="--steps 2 --warmup 3 --no-cpu --no-cov-legs"
RX='regex:tnf_|cov_|clr_|wmedian|logratio|select_samples'
timeout 900 python bench.py $ARGS > gpurun_out/plain_$tag.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" -c 900 --csv --log-file gpurun_out/launches_$tag.csv python bench.py $ARGS > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'^tnf_kernel|cov_stream2_kernel' -s 4 -c 2 -o gpurun_out/prof_full_$tag python bench.py $ARGS > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log | cut -c1-160
