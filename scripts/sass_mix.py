#!/usr/bin/env python
"""Instruction mix of the loops of one kernel in a .o/.so: python scripts/sass_mix.py <obj> <mangled-name-substring>
Prints, for every backward branch (loop), the number of SASS instructions in the loop body by issue pipe.
Pipe classes follow B300_MICROARCH.md: alu = IADD3/LOP3/SHF/PRMT/SEL/ISETP/VIMNMX/VIADDMNMX/LEA...; fma = IMAD*/FFMA...;
lsu = LD*/ST*/ATOM*/RED*; other = SHFL, REDUX, BAR, BRA, uniform-datapath ops (U*), S2R ..."""
import re
import subprocess
import sys

ALU = ("IADD3", "IADD", "LOP3", "SHF", "PRMT", "SEL", "ISETP", "VIMNMX", "VIADDMNMX", "LEA", "VIADD", "PLOP3", "IABS", "FLO", "POPC",
       "BREV", "MOV", "VABSDIFF", "IMNMX", "CS2R", "P2R", "R2P", "FMNMX", "FSEL", "FSETP", "SGXT", "BMSK", "LOP", "VIMNMX3")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2", "DFMA", "DADD", "DMUL")
LSU = ("LDG", "STG", "LDS", "STS", "LD", "ST", "ATOM", "ATOMS", "ATOMG", "RED", "LDC", "LDSM", "LDGSTS", "LDL", "STL")


def classify(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UTMALDG",):
        return "uniform"
    if base in FMA:
        return "fma"
    if base in ALU:
        return "alu"
    if base in LSU:
        return "lsu"
    return "other:" + base


def main():
    obj, name = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    for f in funcs[1:]:
        fname = f.split("\n", 1)[0].strip()
        if name not in fname:
            continue
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                addr = int(m.group(1), 16)
                body = m.group(2).strip()
                body = re.sub(r"^@!?U?P\d+\s+", "", body)
                ins.append((addr, body))
        print(f"== {fname}: {len(ins)} instructions")
        for i, (addr, body) in enumerate(ins):
            m = re.match(r"BRA(\.U)?\s+(?:!?U?P\d+,\s*)?(?:`\()?(0x[0-9a-f]+)", body)
            if m and int(m.group(2), 16) < addr:
                lo = int(m.group(2), 16)
                loop = [b for a, b in ins if lo <= a <= addr]
                mix = {}
                for b in loop:
                    c = classify(b.split()[0])
                    mix[c] = mix.get(c, 0) + 1
                main_ = {k: v for k, v in mix.items() if not k.startswith("other")}
                other = {k[6:]: v for k, v in mix.items() if k.startswith("other")}
                print(f"  loop 0x{lo:04x}..0x{addr:04x}: {len(loop):4d} instr  {main_}  other={other}")


if __name__ == "__main__":
    main()
