#!/bin/bash
# One GPU call that times the prepared kernel variants and profiles the two ends.
# Here (no GPU):   bash tools/build_variants.sh
# Then:            gpurun --timeout 1500 -- 'bash tools/time_variants_on_box.sh'
# Results: gpurun_out/tune.jsonl (one JSON line per variant: 1e9 samples/s per shape, the
# rare-path counters, a checksum of the first 1000 ll), gpurun_out/prof_<variant>.ncu-rep.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/tune_clocks.csv
timeout 900 python tools/tune_variants.py > gpurun_out/tune.log 2>&1
for v in base trim_fixed_halley_t2048; do
  [ -f build/variants/$v.so ] && timeout 300 bash tools/profile_variant.sh build/variants/$v.so $v
done
# parity of the leanest variant through the C ABI (the whole GPU suite against that library)
TJB_LIB_PATH=$PWD/build/variants/trim_fixed_halley_t2048.so timeout 900 python -m pytest tests -m gpu -x -q \
  > gpurun_out/tune_parity_trim_fixed_halley_t2048.log 2>&1
tail -3 gpurun_out/tune_parity_trim_fixed_halley_t2048.log
cat gpurun_out/tune.jsonl
