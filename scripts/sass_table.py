#!/usr/bin/env python
"""Instruction-mix table of libssm.so's kernels (whole-kernel SASS, no GPU needed):
    python scripts/sass_table.py [lib] [--min-instr N] > profiles/rN_sass_mix.md
Per kernel: registers / spills / static shared memory (cuobjdump -res-usage), SASS instruction count, instructions by issue pipe
(classes of scripts/sass_mix.py, after B300_MICROARCH.md) and the counts of the mnemonics the design leans on: packed 16x2
min/max (VIMNMX3, VIADDMNMX = the DPX family, VIMNMX), byte permutes (PRMT), integer multiply-add used as an adder on the FMA pipe
(IMAD), warp reductions (REDUX, CREDUX), shuffles, TMA bulk copies / prefetches (UBLKCP, UBLKPF), distributed-shared-memory stores
(STAS), mbarrier operations (SYNCS), cluster barriers (UCGABAR), atomics / reductions (ATOM*, RED*)."""
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_mix import classify  # noqa: E402

KEYS = ["VIMNMX3", "VIADDMNMX", "VIMNMX", "PRMT", "IMAD", "LOP3", "SHF", "LEA", "REDUX", "CREDUX", "SHFL", "LDS", "STS", "LDG", "STG", "UBLKCP", "UBLKPF",
        "STAS", "SYNCS", "UCGABAR", "BAR", "ATOM", "RED", "MATCH", "DADD", "DMUL"]


def main():
    argv = sys.argv[1:]
    min_instr = 400
    if "--min-instr" in argv:
        i = argv.index("--min-instr")
        min_instr = int(argv[i + 1])
        del argv[i:i + 2]
    lib = argv[0] if argv else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "semantic_slam_mapping_b200", "libssm.so")
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.split("\n"):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    rows = []
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        fname = f.split("\n", 1)[0].strip()
        ops = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
            if m:
                body = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
                ops.append(body.split()[0])
        if len(ops) < min_instr:
            continue
        pipes = {"alu": 0, "fma": 0, "lsu": 0, "uniform": 0, "other": 0}
        for o in ops:
            c = classify(o)
            pipes[c if c in pipes else "other"] += 1
        counts = {k: 0 for k in KEYS}
        for o in ops:
            base = o.split(".")[0]
            for k in KEYS:
                if base == k or (k in ("ATOM", "RED") and base.startswith(k) and base != "REDUX") or (k == "MATCH" and base == "MATCH"):
                    counts[k] += 1
                    break
        demangled = subprocess.run(["c++filt", fname], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", demangled).replace("ssm::", "").replace("(anonymous namespace)::", "")
        short = re.sub(r"^void ", "", short)
        rows.append((short, usage.get(fname, (0, 0, 0)), len(ops), pipes, counts))
    rows.sort(key=lambda r: -r[2])
    print(f"# SASS instruction mix of {os.path.basename(lib)} ({', '.join(archs)}; kernels of >= {min_instr} instructions; scripts/sass_table.py)\n")
    hdr = ["kernel", "regs", "smem B", "local B", "instr", "alu", "fma", "lsu", "uniform", "other"] + KEYS
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    for short, (reg, sh, loc), n, pipes, counts in rows:
        cells = [f"`{short}`", reg, sh, loc, n, pipes["alu"], pipes["fma"], pipes["lsu"], pipes["uniform"], pipes["other"]] + [counts[k] for k in KEYS]
        print("| " + " | ".join(str(c) for c in cells) + " |")
    tot = {k: sum(r[4][k] for r in rows) for k in KEYS}
    print(f"\nTotals over these kernels: " + ", ".join(f"{k} {v}" for k, v in tot.items() if v))
    print("No tensor-core instructions (UTC*MMA / HMMA / LDTM): " + str(not re.search(r"UTC\w*MMA|HMMA|LDTM", sass)))


if __name__ == "__main__":
    main()
