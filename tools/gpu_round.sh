#!/bin/bash
# One GPU call for a round of measurements (run under gpurun from the repo root):
#   1. pytest -m gpu against the shipped library
#   2. kernel variants timed (tools/tune_variants.py -> gpurun_out/tune.jsonl)
#   3. bench.py (default arguments) -> gpurun_out/bench_g1.json
#   4. ncu launch list, counters and one full capture of the shipped kernel (tools/profile_gpu.sh)
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/clocks_idle.csv
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
if ls build/variants/*.so > /dev/null 2>&1; then
  timeout 600 python tools/tune_variants.py > gpurun_out/tune.log 2>&1
fi
timeout 900 python bench.py > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err
tail -c 600 gpurun_out/bench_g1.json
timeout 900 bash tools/profile_gpu.sh > gpurun_out/profile.log 2>&1
ls -la gpurun_out
