#!/bin/bash
# One GPU call that times the prepared kernel variants and profiles the two ends.
# Here (no GPU):   bash tools/build_variants.sh
# Then:            gpurun --timeout 1500 -- 'bash tools/time_variants_on_box.sh [profiled variants...]'
# Results: gpurun_out/tune.jsonl (one JSON line per variant: 1e9 samples/s per shape, the
# rare-path counters, a checksum of the first 1000 ll), gpurun_out/prof_<variant>.ncu-rep,
# and the GPU parity suite run against the first profiled variant.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PROF=${@:-shipped}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/tune_clocks.csv
timeout 1200 python tools/tune_variants.py > gpurun_out/tune.log 2>&1
for v in $PROF; do
  [ -f build/variants/$v.so ] && timeout 300 bash tools/profile_variant.sh build/variants/$v.so $v
done
first=$(echo $PROF | cut -d' ' -f1)
# parity through the C ABI: the whole GPU suite against that library
TJB_LIB_PATH=$PWD/build/variants/$first.so timeout 900 python -m pytest tests -m gpu -x -q \
  > gpurun_out/tune_parity_$first.log 2>&1
tail -3 gpurun_out/tune_parity_$first.log
cat gpurun_out/tune.jsonl
