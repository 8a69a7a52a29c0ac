#!/bin/bash
# composition parity tests on a build whose tnf_kernel checks every load address and computed index (traps on a violation),
# then the normal build again.  compute-sanitizer is not available on this pool.
make -C magpurify2_b200/csrc -B EXTRA="-DMP2_TNF_BOUNDS=1" mp2_tnf.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1 || echo BUILD FAILED
echo "bounds-checked build: $(timeout 900 python -m pytest tests/test_tnf_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -1)"
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --samples 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bounds-checked build, C3 bench: kernel_ms', round(d['roofline']['kernel_ms'],4), 'packed identical', d['packed_2bit_path']['counts_identical_to_ascii_path'])"
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --samples 0 --long-contigs --mags 400 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bounds-checked build, C5 bench: kernel_ms', round(d['roofline']['kernel_ms'],4), 'packed identical', d['packed_2bit_path']['counts_identical_to_ascii_path'])"
make -C magpurify2_b200/csrc -B mp2_tnf.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1
