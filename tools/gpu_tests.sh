#!/bin/bash
# GPU-box visit: all GPU tests (no -x: every failure is wanted), smoke, and a short bench of both arms.
# usage: bash tools/gpu_tests.sh [pytest args]
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; nproc >> gpurun_out/box.txt
timeout 1500 python -m pytest tests -m gpu -q "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 4 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -5; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -40; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
