#!/bin/bash
# usage: gpu_r2_score.sh "<EXTRA flags variant 1>" ...  -- per variant: rebuild mp2_score.o, scoring + pipeline parity, C3 step and per-kernel times
for v in "$@"; do
  echo "=== variant: $v"
  make -C magpurify2_b200/csrc -B EXTRA="$v" mp2_score.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1 || echo BUILD FAILED
  timeout 600 python -m pytest tests/test_score_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -1
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --samples 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step', d['ms_per_step_min_median_max'], 'value', round(d['value'],1), d['kernel_ms_per_step'])"
done
