#!/bin/bash
# Round-2 GPU-box visit: [tests,] smoke, both bench arms, ncu launch lists of one forward+loss window and one training step of the bench workload.  usage: bash tools/gpu_r02.sh <tag> [notests]
set -u
TAG=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; nproc >> gpurun_out/box.txt
if [ "${2:-}" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench.err; echo "bench rc=$?"
for M in fwd train; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_${M}_$TAG.csv python tools/profile_step.py $M > /dev/null 2>&1; echo "ncu $M rc=$?"
  python tools/summarize_launches.py gpurun_out/launches_${M}_$TAG.csv > gpurun_out/launches_${M}_$TAG.txt 2>&1
done
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$TAG.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_$TAG.log | head -30
tail -2 gpurun_out/smoke.log; tail -5 gpurun_out/bench.err
head -30 gpurun_out/launches_fwd_$TAG.txt; head -40 gpurun_out/launches_train_$TAG.txt
cat gpurun_out/bench_$TAG.json | cut -c1-6000; cat gpurun_out/bench_ref_$TAG.json | cut -c1-400
