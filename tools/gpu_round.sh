#!/bin/bash
# One GPU-box visit: [tests,] smoke, bench (both arms), ncu launch list of the bench command, full ncu capture of the top kernel.
# usage: bash tools/gpu_round.sh [tag] [notests]
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; nproc >> gpurun_out/box.txt
if [ "${2:-}" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lif_conv_fwd_tc -s 2 -c 3 -o gpurun_out/prof_tc_fwd_$TAG python tools/run_tc_once.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json
