#!/bin/bash
# build-parameter sweep of cov_stream_kernel on the GPU box: L2 prefetch distance (events ahead)
run() { python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['coverage']; print('$1', 'kernel_ms', round(c['roofline']['kernel_ms'],4), 'ms/step', round(c['ms_per_step'],2))"; }
for pf in 256 768 1536 384; do
  make -C magpurify2_b200/csrc -B EXTRA=-DMP2_C3_PF=$pf mp2_cov.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1
  run "prefetch=$pf"
done
