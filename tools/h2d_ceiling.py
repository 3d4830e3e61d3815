"""Host -> device copy ceiling of the box, for the e2e scaling question:
how fast can N ranks pull page-locked host memory into their GPUs when *nothing else*
runs -- no kernel, no library code, plain cudaMemcpyAsync from pinned buffers -- alone
(H2D only) and in the e2e proportion (32 B in : 8 B out per sample)?  If the aggregate
of this bare copy loop stops scaling with the number of GPUs at the rate the e2e path
reaches, the limit is the host (DRAM / PCIe root / IO die), not the staging code.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/h2d_ceiling.py            (or plain python for N = 1)

Rank 0 prints one JSON line: aggregate GB/s per mode, per-rank min / max, host memcpy
bandwidth of one thread per rank running concurrently, and the NUMA / affinity facts.
"""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = 1 << 27  # doubles: 1 GiB per buffer
    src = torch.empty(n, dtype=torch.float64).pin_memory()
    src.uniform_()
    dst_h = torch.empty(n // 4, dtype=torch.float64).pin_memory()
    dev = torch.empty(n, dtype=torch.float64, device="cuda")
    dev_out = torch.empty(n // 4, dtype=torch.float64, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(mode, reps=6):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s_in):
                    dev.copy_(src, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s_out):
                    dst_h.copy_(dev_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        b_in = reps * n * 8 if mode in ("h2d", "both") else 0
        b_out = reps * (n // 4) * 8 if mode in ("d2h", "both") else 0
        return b_in / dt / 1e9, b_out / dt / 1e9

    def host_memcpy(reps=3):
        a, b = src.numpy(), np.empty(n)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            np.copyto(b, a)
        return reps * n * 8 / (time.perf_counter() - t0) / 1e9

    out = {}
    for mode in ("h2d", "d2h", "both"):
        run(mode, 2)
        gin, gout = run(mode)
        t = torch.tensor([gin, gout, gin, gout, -gin, -gout], dtype=torch.float64, device="cuda")
        if world > 1:
            s = t.clone()
            dist.all_reduce(s, op=dist.ReduceOp.SUM)
            m = t.clone()
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            t = torch.cat([s[:2], m[2:]])
        v = t.tolist()
        out[mode] = {"h2d_GBs_aggregate": v[0], "d2h_GBs_aggregate": v[1], "h2d_GBs_max_rank": v[2],
                     "d2h_GBs_max_rank": v[3], "h2d_GBs_min_rank": -v[4], "d2h_GBs_min_rank": -v[5]}
    hm = torch.tensor([host_memcpy()], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(hm, op=dist.ReduceOp.SUM)
    if rank == 0:
        numa = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")) \
            if os.path.isdir("/sys/devices/system/node") else None
        print(json.dumps({"n_gpus": world, "buffer_GiB": n * 8 / 2**30, "copies": out,
                          "host_memcpy_GBs_aggregate_one_thread_per_rank": float(hm.item()),
                          "cpus_allowed": len(os.sched_getaffinity(0)), "numa_nodes": numa,
                          "what": "bare cudaMemcpyAsync from / to page-locked buffers, no kernels"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
