#!/bin/bash
# threads per weighted-median CTA (rebuilds mp2_score.o on the box)
for w in 64 128 512 256; do
  if make -C magpurify2_b200/csrc -B EXTRA=-DMP2_WM_THREADS=$w mp2_score.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1; then
    python -m pytest tests/test_score_gpu.py -m gpu -x -q 2>&1 | tail -1
    python bench.py --steps 5 --warmup 3 --no-cpu --samples 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('wm_threads=$w', 'wmedian ms/step', d['kernel_ms_per_step']['wmedian_kernel'], 'step', round(d['ms_per_step'],3))"
  else echo "build failed for $w"; fi
done
