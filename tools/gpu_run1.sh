#!/bin/bash
# GPU run 1: parity tests, smem-atomics microbenchmark, first bench line, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
./tools/ubench_hist > gpurun_out/ubench_hist.txt 2>&1; cat gpurun_out/ubench_hist.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err; echo "bench exit $?"; cat gpurun_out/bench_r1_a.json; tail -3 gpurun_out/bench_r1_a.err
python bench.py --steps 3 --warmup 3 --long-contigs --mags 400 --no-cpu > gpurun_out/bench_r1_c5.json 2>> gpurun_out/bench_r1_a.err; cat gpurun_out/bench_r1_c5.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2>&1; cat gpurun_out/bench_r1_ref.json
python bench.py --steps 2 --warmup 1 --mags 200 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --mags 200 --no-cpu > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 2 --warmup 1 --mags 200 --no-cpu > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tnf_kernel -s 1 -c 2 -o gpurun_out/prof_tnf_r1 python bench.py --steps 2 --warmup 1 --mags 200 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
