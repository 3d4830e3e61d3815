#!/bin/bash
# Builds the kernel tuning variants into build/variants/*.so (git-ignored, shipped to the
# GPU box by gpurun).  On the box: python tools/tune_variants.py  -> gpurun_out/tune.jsonl.
# Each variant's numerics are checked on the CPU by
# tests/test_host_logic.py::test_host_emulated_tuning_variants; static instruction counts of
# the epoch loop's main path: python tools/sass_main_path.py build/variants/<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
build() {  # name, flags...
  local name=$1; shift
  nvcc $F "$@" thejoker_b200/csrc/tjb_api.cu -o build/variants/$name.so &
}
build base
build trim -DTJB_TRIM=1
build fixed -DTJB_PHASE_FIXED=1
build trim_fixed -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1
wait
build trim_fixed_e3 -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1 -DTJB_EPOCHS_PER_ITER=3
build trim_fixed_e4 -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1 -DTJB_EPOCHS_PER_ITER=4
build trim_fixed_192x3 -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1 -DTJB_LL_THREADS=192 -DTJB_LL_MIN_CTAS=3
build trim_fixed_t2048 -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1 -DTJB_TRIG_TABLE_LOG2=11
wait
build trim_fixed_halley -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1 -DTJB_HALLEY=1
build trim_fixed_halley_e3 -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1 -DTJB_HALLEY=1 -DTJB_EPOCHS_PER_ITER=3
build trim_fixed_halley_16 -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1 -DTJB_HALLEY=1 -DTJB_NEED_LOG2=16
build trim_fixed_halley_18 -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1 -DTJB_HALLEY=1 -DTJB_NEED_LOG2=18
wait
build trim_fixed_halley_t2048 -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1 -DTJB_HALLEY=1 -DTJB_TRIG_TABLE_LOG2=11
build trim_fixed_halley_t2048_e3 -DTJB_TRIM=1 -DTJB_PHASE_FIXED=1 -DTJB_HALLEY=1 -DTJB_TRIG_TABLE_LOG2=11 -DTJB_EPOCHS_PER_ITER=3
wait
ls -la build/variants
