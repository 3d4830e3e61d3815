"""Static instruction count of the MAIN path of the likelihood kernel's epoch loop.

usage: python tools/sass_main_path.py <lib.so> [kernel-substring]

The epoch loop of marginal_ll_kernel<L,false> is the largest back-edge loop that holds
2 x kEpochsPerIter MUFU.SIN; most of its static body is the rare path (lanes that need
extra Householder passes, unrolled by ptxas), so an opcode histogram of the whole body
says little.  This walks the loop the way a converged warp executes it: from the loop
head, falling through predicated forward branches (the rare-path entries, all guarded by
a VOTE) and following unconditional ones, until the back edge.  Prints the opcode
histogram of that trace and the FP64 / other split per epoch.  No GPU needed.
"""
import re
import subprocess
import sys
from collections import Counter


def kernel_sass(lib, pat):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    ins, on = [], False
    for line in out.splitlines():
        if "Function :" in line:
            on = pat in line
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def main_path(ins, epochs_per_iter):
    addr = {a: i for i, (a, _) in enumerate(ins)}
    # loop candidates: backward branches whose body holds 2 * epochs MUFU.SIN
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a:
            continue
        body = ins[addr[tgt]: addr[a] + 1]
        n_sin = sum("MUFU.SIN" in x[1] for x in body)
        if n_sin in (epochs_per_iter, 2 * epochs_per_iter) and (best is None or len(body) > len(best[2])):
            best = (tgt, a, body)
    if best is None:
        raise SystemExit("no epoch loop found")
    head, back, _ = best
    trace, i, guard = [], addr[head], 0
    while True:
        a, t = ins[i]
        trace.append((a, t))
        if a == back:
            break
        guard += 1
        if guard > 5000:
            raise SystemExit("trace did not reach the back edge")
        m = re.match(r"(@!?U?P\d\s+)?BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and not m.group(1) and "BRA.U" not in t and ".DIV" not in t:
            i = addr[int(m.group(2), 16)]  # unconditional: follow
            continue
        i += 1  # predicated / divergence-check branches: not taken on the main path
    return head, back, trace


def main():
    lib = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else "marginal_ll_kernelILi2ELb0E"
    epi = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ins = kernel_sass(lib, pat)
    head, back, trace = main_path(ins, epi)
    c = Counter()
    for _, t in trace:
        t = re.sub(r"^@!?U?P\d\s+", "", t)
        c[t.split()[0].split(".")[0]] += 1
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    n = len(trace)
    print(f"{lib} {pat}: loop 0x{head:x}-0x{back:x}, main path {n} instructions per {epi} epochs")
    print(f"  per epoch: {fp64 / epi:.1f} FP64 + {(n - fp64) / epi:.1f} other = {n / epi:.1f}")
    print("  " + ", ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda x: -x[1])))


if __name__ == "__main__":
    main()
