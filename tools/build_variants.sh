#!/bin/bash
# Builds kernel tuning variants into build/variants/*.so (git-ignored, shipped to the GPU
# box by gpurun).  On the box: python tools/tune_variants.py  -> gpurun_out/tune.jsonl
# (TJB_TUNE_SHAPES=L for the sweep over n_linear and both jitter kernels).
# Each variant's numerics are checked on the CPU by
# tests/test_host_logic.py::test_host_emulated_tuning_variants; static instruction counts of
# the epoch loop's main path: python tools/sass_main_path.py build/variants/<name>.so <kernel> <epochs>
#
# Knobs (kepler.cuh, marginal_ll.cuh; DESIGN.md section 4.1 has what each one measured):
#   TJB_WIDE_THREADS / TJB_WIDE_EPOCHS   shape of the constant-jitter kernels with L <= 4 (0: off)
#   TJB_LL_THREADS / TJB_EPOCHS_PER_ITER shape of all the other likelihood kernels
#   TJB_TRIG2, TJB_FINE_LOG2             two-level trig table and its fine resolution
#   TJB_XZ                               z-stage reciprocal refined from the step's
#   TJB_UCONST, TJB_UROW                 uniform-register constants / parameter-block epoch rows
#   TJB_TRIM, TJB_PHASE_FIXED, TJB_HALLEY, TJB_NEED_LOG2, TJB_TRIG_TABLE(_LOG2)   round-2a/b switches
# usage: bash tools/build_variants.sh            (the default set below)
#        bash tools/build_variants.sh name "-DFLAG=.. -DFLAG=.." [name flags ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
build() {  # name, flags...
  local name=$1; shift
  nvcc $F "$@" thejoker_b200/csrc/tjb_api.cu -o build/variants/$name.so &
}
if [ $# -ge 2 ]; then
  while [ $# -ge 2 ]; do build "$1" $2; shift 2; [ $(jobs -r | wc -l) -ge 4 ] && wait; done
  wait
else
  rm -f build/variants/*.so
  build shipped
  build one_shape -DTJB_WIDE_THREADS=0
  build wide_e3 -DTJB_WIDE_EPOCHS=3
  build no_trig2 -DTJB_TRIG2=0
  wait
  build xz -DTJB_XZ=1
  build smem_rows -DTJB_UROW=0
  build no_uconst -DTJB_UCONST=0
  build r02b_loop -DTJB_WIDE_THREADS=0 -DTJB_LL_THREADS=256 -DTJB_LL_MIN_CTAS=2 -DTJB_EPOCHS_PER_ITER=3 -DTJB_TRIG2=0 -DTJB_UCONST=0 -DTJB_UROW=0
  wait
fi
ls -la build/variants
