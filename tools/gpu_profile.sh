#!/bin/bash
# GPU-box visit for evidence: bench (both arms), ncu launch list of the bench command, ncu --set full of the dominant kernels.
# usage: bash tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; nproc >> gpurun_out/box.txt
timeout 300 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/launches_bench_$TAG.csv > gpurun_out/launches_bench_$TAG.txt 2>&1
head -40 gpurun_out/launches_bench_$TAG.txt
cat gpurun_out/bench_$TAG.json | cut -c1-1500; cat gpurun_out/bench_ref_$TAG.json | cut -c1-600
