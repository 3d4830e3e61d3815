#!/bin/bash
# One multi-GPU call (gpurun --gpus N -- 'bash tools/gpu_round_scale.sh N'): the multi-GPU tests,
# bench.py at 2 / 4 / 8 ranks (those that fit N) and BASELINE configs 3 and 4 at N ranks.
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_g$N.log 2>&1
tail -2 gpurun_out/pytest_multi_g$N.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 \
      > gpurun_out/bench_g$n.json 2> gpurun_out/bench_g$n.err
    tail -c 300 gpurun_out/bench_g$n.json
  fi
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29511 tools/bench_configs.py 28 > gpurun_out/configs_g$N.json 2> gpurun_out/configs_g$N.err
tail -c 700 gpurun_out/configs_g$N.json
