#!/bin/bash
# usage: gpu_r2_tnf2.sh "<EXTRA flags variant 1>" "<variant 2>" ...  -- per variant: rebuild mp2_tnf.o, composition parity, C3 + C5 kernel time
mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --samples 0 $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'step', d['ms_per_step_min_median_max'], 'packed', round(d['packed_2bit_path']['count_ms'],3), d['packed_2bit_path']['counts_identical_to_ascii_path'])"; }
for v in "$@"; do
  echo "=== variant: $v"
  make -C magpurify2_b200/csrc -B EXTRA="$v" mp2_tnf.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1 || echo BUILD FAILED
  timeout 600 python -m pytest tests/test_tnf_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -1
  run "C3"
  run "C3 again"
  run "C5" "--long-contigs --mags 400"
done
