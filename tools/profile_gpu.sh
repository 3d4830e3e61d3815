#!/bin/bash
# Run on the GPU box (under gpurun): launch list, counter pass and one full ncu capture of
# the hot kernel.  Outputs land in gpurun_out/ ; summaries are copied into profiles/ here.
set -x
mkdir -p gpurun_out
# what was profiled: hash of the CUDA sources + flags (bench.py flags counters from other
# sources as stale) and the register count of the hot kernel in the shipped binary
python -c "from thejoker_b200 import _lib; print(_lib.source_hash())" > gpurun_out/source_hash.txt
cuobjdump -res-usage ${TJB_LIB_PATH:-thejoker_b200/libthejoker_b200.so} 2>/dev/null \
  | grep -A1 "marginal_ll_kernelILi2ELb0ENS_9PriorViewENS_14EpochRowsParam" | grep -o "REG:[0-9]*" > gpurun_out/hot_kernel_regs.txt
BENCH="python bench.py --steps 2 --warmup 3 --log2-e2e 22 --no-cpu-baseline"
# 1. every launch with its device time (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches_bench.json 2> gpurun_out/launches.err
# 2. instruction / DRAM counters of the hot kernel at the bench size
ncu --clock-control none -k regex:marginal_ll_kernel -s 3 -c 1 --csv \
    --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__occupancy_limit_registers,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_uniform.sum \
    --log-file gpurun_out/counters.csv $BENCH > /dev/null 2> gpurun_out/counters.err
# 3. full capture (source-level stalls) on a smaller launch to keep replay time short
ncu --set full --clock-control none --import-source on -k regex:marginal_ll_kernel -s 3 -c 1 \
    -f -o gpurun_out/prof_ll python bench.py --steps 2 --warmup 3 --log2-prior 25 --log2-e2e 20 --no-cpu-baseline \
    > /dev/null 2> gpurun_out/prof.err
ls -la gpurun_out
