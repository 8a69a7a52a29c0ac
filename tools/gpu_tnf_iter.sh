#!/bin/bash
# usage: gpu_tnf_iter.sh <tag>  -- parity tests (both counting modes), C3 + C5 bench, one full ncu capture
tag=${1:-x}
mkdir -p gpurun_out
for ms in "5 14" "5 15" "4 14"; do
  set -- $ms
  MP2_TNF_MODE=$1 MP2_TNF_PIECE_SHIFT=$2 python -m pytest tests/test_tnf_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -3
done
MP2_TNF_PIECE_SHIFT=14 python bench.py --steps 5 --warmup 3 --no-cpu --samples 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('C3 16KiB pieces kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3))"
python bench.py --steps 5 --warmup 3 --samples 0 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --steps 5 --warmup 3 --no-cpu --samples 0 --long-contigs --mags 400 > gpurun_out/bench_${tag}_c5.json 2>> gpurun_out/bench_$tag.err
python - <<PY
import json
for f,n in (('gpurun_out/bench_$tag.json','C3'),('gpurun_out/bench_${tag}_c5.json','C5')):
    try:
        d=json.load(open(f))
        print(n, d.get('cpu_baseline',{}).get('parity_with_gpu'), round(d['value'],1),'Gbp/s kernel_ms',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],1), 'packed', d.get('packed_2bit_path',{}).get('count_ms'))
    except Exception as e: print('fail',n,e); print(open('gpurun_out/bench_$tag.err').read()[-1500:])
PY
if [ -z "$NO_NCU" ]; then
ncu --set full --clock-control none --import-source on -k regex:'^tnf_kernel' -s 1 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 1 --no-cpu --samples 0 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-200
fi
