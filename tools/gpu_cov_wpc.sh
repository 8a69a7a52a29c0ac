#!/bin/bash
# warps per CTA of cov_stream_kernel (rebuilds mp2_cov.o on the box)
run() { python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['coverage']; print('$1', 'kernel_ms', round(c['roofline']['kernel_ms'],4), 'ms/step', round(c['ms_per_step'],2))"; }
for w in 2 4 1; do
  make -C magpurify2_b200/csrc -B EXTRA=-DMP2_C3_WPC=$w mp2_cov.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1
  python -m pytest tests/test_cov_gpu.py -m gpu -x -q 2>&1 | tail -1
  run "warps_per_cta=$w"
done
