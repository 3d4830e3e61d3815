#!/bin/bash
# One multi-GPU call (gpurun --gpus N -- 'bash tools/gpu_round_multi.sh N'): the multi-GPU
# tests the 1-GPU driver box skips, bench.py under torchrun (accept parity across ranks),
# the multi-star batch on 1 and N GPUs, and the bare host->device copy ceiling.
N=${1:-2}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo_g$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/pytest_multi_g$N.log 2>&1
tail -5 gpurun_out/pytest_multi_g$N.log
timeout 600 $TR --master-port 29501 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
tail -c 1500 gpurun_out/bench_g$N.json
timeout 300 python tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_g1.json 2> gpurun_out/h2d_ceiling_g1.err
timeout 300 $TR --master-port 29502 tools/h2d_ceiling.py > gpurun_out/h2d_ceiling_g$N.json 2> gpurun_out/h2d_ceiling_g$N.err
cat gpurun_out/h2d_ceiling_g1.json gpurun_out/h2d_ceiling_g$N.json
if [ "$N" -le 2 ]; then
  timeout 600 python tools/bench_multistar.py 1024 22 4 > gpurun_out/multistar_g1.json 2> gpurun_out/multistar_g1.err
  timeout 600 $TR --master-port 29503 tools/bench_multistar.py 2048 22 4 > gpurun_out/multistar_g$N.json 2> gpurun_out/multistar_g$N.err
  tail -3 gpurun_out/multistar_g1.json gpurun_out/multistar_g$N.json
fi
ls -la gpurun_out
