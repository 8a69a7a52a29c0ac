#!/bin/bash
mkdir -p gpurun_out
for m in 5 7; do
  MP2_TNF_MODE=$m python -m pytest tests/test_tnf_gpu.py -m gpu -x -q 2>&1 | tail -3
  MP2_TNF_MODE=$m python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_mode$m.json 2> gpurun_out/bench_mode$m.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_mode$m.json'))
    print('mode $m C3', round(d['value'],1),'Gbp/s kernel_ms',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],1))
except Exception as e: print('fail',e); print(open('gpurun_out/bench_mode$m.err').read()[-2000:])
PY
  MP2_TNF_MODE=$m python bench.py --steps 5 --warmup 3 --no-cpu --long-contigs --mags 400 > gpurun_out/bench_mode${m}_c5.json 2>> gpurun_out/bench_mode$m.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_mode${m}_c5.json'))
    print('mode $m C5', round(d['value'],1),'Gbp/s kernel_ms',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3))
except Exception as e: print('fail',e)
PY
done
