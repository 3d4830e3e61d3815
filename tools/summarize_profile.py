"""Turn the ncu outputs in gpurun_out/ (tools/profile_gpu.sh) into the tracked
summaries under profiles/: launch list shares, instruction / DRAM counters of the hot
kernel, stall reasons, and profiles/kernel_counts.json (read by bench.py for the
executed-flop roofline fraction and the measured DRAM traffic per launch).

usage: python tools/summarize_profile.py r01a
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = {"tag": tag}
md = [f"# ncu summary {tag}", ""]
for name, key in (("source_hash.txt", "source_sha256"), ("hot_kernel_regs.txt", "cuobjdump_hot_kernel")):
    f = os.path.join(G, name)
    if os.path.exists(f):
        out[key] = open(f).read().strip()
md += [f"* source hash of the profiled build (`_lib.source_hash()`): `{out.get('source_sha256')}`",
       f"* `cuobjdump -res-usage` of `marginal_ll_kernel<2,false,PriorView,EpochRowsParam>` in that library: "
       f"`{out.get('cuobjdump_hot_kernel')}`", ""]

# ---- launch list --------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 5]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.OrderedDict()
for r in rows[1:]:
    k = r[ik]
    k = k[:k.index("(")] if "(" in k else k
    tot.setdefault(k, [0, 0.0])
    tot[k][0] += 1
    tot[k][1] += float(r[iv].replace(",", ""))
total = sum(v[1] for v in tot.values())
md += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none` over "
       "`python bench.py --steps 2 --warmup 3 --log2-e2e 22 --no-cpu-baseline`; cold-cache, "
       "serialised: shares, not absolutes)", "", "| launches | total ms | share | kernel |",
       "|---:|---:|---:|---|"]
for k, v in sorted(tot.items(), key=lambda x: -x[1][1]):
    md.append(f"| {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / total:.1f}% | `{k[-110:]}` |")
out["launch_share_marginal_ll"] = sum(v[1] for k, v in tot.items() if "marginal_ll_kernel" in k) / total
md.append("")

# ---- counters -------------------------------------------------------------------
rows = [r for r in csv.reader(open(os.path.join(G, "counters.csv"))) if len(r) > 5]
hdr = rows[0]
iname, ival, iunit = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
c = {r[iname]: float(r[ival].replace(",", "")) for r in rows[1:]}
grid = c.get("launch__grid_size")
n = 1 << 28
dfma = c["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
dmul = c["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
dadd = c["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
flop = 2 * dfma + dmul + dadd
t_s = c["gpu__time_duration.sum"] * 1e-9
out.update(
    n_samples=n, n_epochs=64,
    executed_fp64_flop_per_sample=flop / n,
    executed_fp64_inst_per_sample_epoch=(dfma + dmul + dadd) / n / 64,
    dram_bytes_per_launch_at_2p28=c["dram__bytes_read.sum"] + c["dram__bytes_write.sum"],
    dram_bytes_read_per_sample=c["dram__bytes_read.sum"] / n,
    dram_bytes_write_per_sample=c["dram__bytes_write.sum"] / n,
    ncu_kernel_ms=t_s * 1e3,
    executed_fp64_tflops_under_ncu=flop / t_s / 1e12,
    fp64_pipe_pct=c.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    issue_active_pct=c.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    warps_active_pct=c.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
    registers_per_thread=c.get("launch__registers_per_thread"), grid=grid,
)
md += ["## Hot kernel counters (`marginal_ll_kernel<2,false>`, 2^28 samples, N=64)", "",
       "| metric | value |", "|---|---:|"]
for k, v in c.items():
    md.append(f"| `{k}` | {v:,.6g} |")
md += ["", f"* executed FP64 flop / sample = (2·DFMA + DMUL + DADD)/2^28 = **{flop / n:,.0f}** "
       f"({(dfma + dmul + dadd) / n / 64:.1f} FP64 instructions per (sample, epoch)); work model "
       "W_sample = 18 988",
       f"* DRAM traffic per launch = {out['dram_bytes_per_launch_at_2p28'] / 1e9:.3f} GB "
       f"= {out['dram_bytes_read_per_sample']:.2f} B read + {out['dram_bytes_write_per_sample']:.2f} B "
       "written per sample (algorithmic: 32 + 8)", ""]

# ---- stall reasons from the full capture ------------------------------------------
rep = os.path.join(G, "prof_ll.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    if len(r) >= 3:
        h, v = r[0], r[2]
        stalls = {}
        for a, b in zip(h, v):
            if "issue_stalled" in a and a.endswith("per_issue_active.ratio"):
                name = a.split("issue_stalled_")[1].split("_per_issue")[0]
                stalls[name] = float(b)
        out["stall_per_issue"] = stalls
        md += ["## Warp stall reasons per issued instruction (`ncu --set full`, 2^25-sample launch)",
               "", "| reason | warps stalled per issue |", "|---|---:|"]
        for k, x in sorted(stalls.items(), key=lambda kv: -kv[1]):
            md.append(f"| {k} | {x:.3f} |")
        md.append("")
        # how busy each unit is: the kernel is spread over four of them
        pipes = collections.OrderedDict()
        for label, key in (
                ("FP64 pipe", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                ("issue slots", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                ("XU pipe (MUFU, F2F)", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                ("shared-memory wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                ("FMA pipe (FP32, IMAD)", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                ("ALU pipe", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                ("LSU instructions", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                ("uniform pipe", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active")):
            if key in h:
                pipes[label] = float(v[h.index(key)])
        for label, key in (("shared-memory wavefronts", "memory_l1_wavefronts_shared"),
                           ("shared-memory wavefronts without bank conflicts", "memory_l1_wavefronts_shared_ideal")):
            if key in h:
                out["smem_" + key.split("memory_l1_")[1]] = float(v[h.index(key)])
        out["pipe_pct"] = pipes
        md += ["## Unit utilisation (same capture, % of peak)", "", "| unit | % |", "|---|---:|"]
        md += [f"| {k} | {x:.1f} |" for k, x in pipes.items()]
        if "smem_wavefronts_shared" in out:
            md += ["", f"* shared-memory wavefronts {out['smem_wavefronts_shared']:.4g} "
                   f"(conflict-free: {out.get('smem_wavefronts_shared_ideal', float('nan')):.4g})"]
        md.append("")

open(os.path.join(P, f"{tag}_ncu_summary.md"), "w").write("\n".join(md) + "\n")
json.dump(out, open(os.path.join(P, f"{tag}_kernel_counts.json"), "w"), indent=1)
json.dump(out, open(os.path.join(P, "kernel_counts.json"), "w"), indent=1)
for f in ("launches.csv", "counters.csv"):
    src = os.path.join(G, f)
    if os.path.exists(src):
        open(os.path.join(P, f"{tag}_{f}"), "w").write(open(src).read())
print("\n".join(md))
