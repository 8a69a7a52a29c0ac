#!/bin/bash
# usage: gpu_r2_tnf.sh <tag> -- composition parity (both flush variants) + C3 / C5 kernel time with and without the bulk-copy clear
tag=${1:-x}
mkdir -p gpurun_out
run() { timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --samples 0 $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'step', d['ms_per_step_min_median_max'], 'packed', round(d['packed_2bit_path']['count_ms'],3), d['packed_2bit_path']['counts_identical_to_ascii_path'])"; }
timeout 600 python -m pytest tests/test_tnf_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -2
run "tma-zero C3"
run "tma-zero C5" "--long-contigs --mags 400"
make -C magpurify2_b200/csrc -B EXTRA="-DMP2_TNF_TMAZERO=0" mp2_tnf.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1
timeout 600 python -m pytest tests/test_tnf_gpu.py -m gpu -x -q 2>&1 | tail -1
run "sts-zero C3"
run "sts-zero C5" "--long-contigs --mags 400"
