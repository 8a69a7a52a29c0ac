#!/bin/bash
# run ordering A/B: coverage parity tests, then per-sample kernel-call times with runs drawn heaviest first / in array order
timeout 900 python -m pytest tests/test_cov_gpu.py -m gpu -x -q 2>&1 | tail -2
echo "--- heaviest first"; timeout 600 python tools/cov_per_sample.py 10
echo "--- array order"; MP2_COV_RUN_ORDER=0 timeout 600 python tools/cov_per_sample.py 10
