#!/bin/bash
# usage: tools/profile_variant.sh <lib.so> <tag>   (on the GPU box)
export TJB_LIB_PATH=$1
ncu --set full --clock-control none --import-source on -k regex:marginal_ll_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_$2 python bench.py --steps 2 --warmup 3 --log2-prior 25 --log2-e2e 20 --no-cpu-baseline \
    > /dev/null 2> gpurun_out/prof_$2.err
