#!/bin/bash
python - <<PY
import torch, time, numpy as np
x=torch.empty(3_000_000_000,dtype=torch.uint8,pin_memory=True); d=torch.empty_like(x,device='cuda')
for _ in range(2):
    torch.cuda.synchronize(); t=time.perf_counter(); d.copy_(x,non_blocking=True); torch.cuda.synchronize(); print('H2D GB/s', 3/(time.perf_counter()-t))
PY
for mb in 16 64 256; do
MP2_SLICE_MB=$mb python bench.py --no-cpu --samples 0 --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slice $mb MB e2e', round(d['e2e']['value'],1))"
done
