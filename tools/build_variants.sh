#!/bin/bash
# Builds kernel tuning variants into build/variants/*.so (git-ignored, shipped to the GPU
# box by gpurun).  On the box: python tools/tune_variants.py  -> gpurun_out/tune.jsonl.
# Each variant's numerics are checked on the CPU by
# tests/test_host_logic.py::test_host_emulated_tuning_variants; static instruction counts of
# the epoch loop's main path: python tools/sass_main_path.py build/variants/<name>.so
# Round-2c set: uniform-register operands (TJB_UCONST: multiplier constants from the
# constant bank, TJB_UROW: epoch rows as a kernel parameter), the two-level trig table
# (TJB_TRIG2), epochs per iteration and CTA shapes on top.  "prev" = the library of the
# previous commit, when its sources are given in $TJB_PREV_SRC (a checkout of HEAD).
set -e
cd "$(dirname "$0")/.."
rm -rf build/variants
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
build() {  # name, flags...
  local name=$1; shift
  nvcc $F "$@" thejoker_b200/csrc/tjb_api.cu -o build/variants/$name.so &
}
if [ -n "$TJB_PREV_SRC" ]; then
  nvcc $F $TJB_PREV_SRC/thejoker_b200/csrc/tjb_api.cu -o build/variants/prev.so &
fi
build base -DTJB_UCONST=0 -DTJB_UROW=0
build uc -DTJB_UCONST=1 -DTJB_UROW=0
build ur -DTJB_UCONST=0 -DTJB_UROW=1
wait
build ucur -DTJB_UCONST=1 -DTJB_UROW=1
build ucur_t2 -DTJB_UCONST=1 -DTJB_UROW=1 -DTJB_TRIG2=1
build ucur_e4 -DTJB_UCONST=1 -DTJB_UROW=1 -DTJB_EPOCHS_PER_ITER=4
build ucur_t2_e4 -DTJB_UCONST=1 -DTJB_UROW=1 -DTJB_TRIG2=1 -DTJB_EPOCHS_PER_ITER=4
wait
build ucur_192x3 -DTJB_UCONST=1 -DTJB_UROW=1 -DTJB_LL_THREADS=192 -DTJB_LL_MIN_CTAS=3
build ucur_t2_192x3 -DTJB_UCONST=1 -DTJB_UROW=1 -DTJB_TRIG2=1 -DTJB_LL_THREADS=192 -DTJB_LL_MIN_CTAS=3
build ucur_e2_128x5 -DTJB_UCONST=1 -DTJB_UROW=1 -DTJB_EPOCHS_PER_ITER=2 -DTJB_LL_THREADS=128 -DTJB_LL_MIN_CTAS=5
build ucur_t2_e2_128x5 -DTJB_UCONST=1 -DTJB_UROW=1 -DTJB_TRIG2=1 -DTJB_EPOCHS_PER_ITER=2 -DTJB_LL_THREADS=128 -DTJB_LL_MIN_CTAS=5
wait
ls -la build/variants
