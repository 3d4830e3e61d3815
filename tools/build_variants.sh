#!/bin/bash
# Builds kernel tuning variants into build/variants/*.so (git-ignored, shipped to the GPU
# box by gpurun).  On the box: python tools/tune_variants.py  -> gpurun_out/tune.jsonl.
# Each variant's numerics are checked on the CPU by
# tests/test_host_logic.py::test_host_emulated_tuning_variants; static instruction counts of
# the epoch loop's main path: python tools/sass_main_path.py build/variants/<name>.so
# Round-2 baseline = the shipped defaults (TJB_TRIM, TJB_PHASE_FIXED, TJB_HALLEY, 2048-node
# table, 3 epochs per iteration); "legacy" is the round-1 loop.
set -e
cd "$(dirname "$0")/.."
rm -rf build/variants
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
build() {  # name, flags...
  local name=$1; shift
  nvcc $F "$@" thejoker_b200/csrc/tjb_api.cu -o build/variants/$name.so &
}
build shipped
build legacy -DTJB_TRIM=0 -DTJB_PHASE_FIXED=0 -DTJB_HALLEY=0 -DTJB_TRIG_TABLE_LOG2=10 -DTJB_EPOCHS_PER_ITER=2
build vote_d2 -DTJB_VOTE_D2=1
build e4 -DTJB_EPOCHS_PER_ITER=4
wait
build vote_d2_e4 -DTJB_VOTE_D2=1 -DTJB_EPOCHS_PER_ITER=4
build need16 -DTJB_NEED_LOG2=16
build t128x4 -DTJB_LL_THREADS=128 -DTJB_LL_MIN_CTAS=4
build e2 -DTJB_EPOCHS_PER_ITER=2
wait
ls -la build/variants
