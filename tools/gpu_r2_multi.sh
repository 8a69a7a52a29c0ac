#!/bin/bash
# usage: gpu_r2_multi.sh <tag> <N> [bench args...] -- bench.py on N GPUs under torchrun (one rank per GPU)
tag=$1; N=$2; shift 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$tag.txt 2>&1
SECONDS=0
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $? after ${SECONDS}s"; tail -3 gpurun_out/bench_$tag.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json")); c=d.get("coverage")
print("N", d["n_gpus"], d["scaling"], "TNF", round(d["value"],1), "Gbp/s ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "ascii", round(d["e2e"]["ascii_upload"]["value"],1), d["host"]["placement_rank0"])
if c: print("COV %.3e reads/s ms %.2f e2e %.3e" % (c["value"], c["ms_per_step"], c["e2e"]["value"]), c["ms_per_step_by_rank"], c["config"].get("resident_event_sets"))
print(d["config"].get("c4"))
PY
