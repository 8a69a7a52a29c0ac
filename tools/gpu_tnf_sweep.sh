#!/bin/bash
# build-parameter sweep of tnf_kernel on the GPU box (rebuilds mp2_tnf.o there): UNROLL x piece shift
run() { python bench.py --steps 5 --warmup 3 --no-cpu --samples 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3))"; }
for u in 2 8 4; do
  make -C magpurify2_b200/csrc -B EXTRA=-DMP2_TNF_UNROLL=$u mp2_tnf.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1
  run "unroll=$u ps=15"
done
MP2_TNF_PIECE_SHIFT=16 run "unroll=4 ps=16"
MP2_TNF_PIECE_SHIFT=14 run "unroll=4 ps=14"
