#!/usr/bin/env python3
"""Instruction mix of the loops of one kernel in libsquava_b200.so (no GPU needed): for every backward branch,
the opcodes between its target and itself, grouped by the pipe that issues them on sm_100 (ALU: LOP3/SEL/ISETP/
SHF/LEA/IADD3/...; FMA-heavy: IMAD* (IMAD.WIDE and IMAD.HI count twice), VIADD as measured in r02).

usage: sass_mix.py <kernel substring> [lib.so]
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else "squava_b200/libsquava_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks:
    if sys.argv[1] not in b.split("\n", 1)[0]:
        continue
    ins = []
    for line in b.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    print(b.split("\n", 1)[0][:100], len(ins), "instructions")
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
            lo = addr[int(m.group(1), 16)]
            body = ins[lo:i + 1]
            c = collections.Counter()
            for _, s in body:
                op = s.split()[1] if s.startswith("@") else s.split()[0]
                c[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("IMAD.") and op.split(".")[1] in ("WIDE", "HI", "SHL", "IADD", "MOV") else "")] += 1
            fma = sum(v * (2 if k in ("IMAD.WIDE", "IMAD.HI") else 1) for k, v in c.items() if k.startswith("IMAD") or k == "VIADD")
            alu = sum(v for k, v in c.items() if k.split(".")[0] in ("LOP3", "SEL", "ISETP", "SHF", "LEA", "IADD3", "PRMT", "POPC", "FLO", "MOV", "IABS", "PLOP3", "VOTE", "BMSK", "SGXT", "LOP"))
            lsu = sum(v for k, v in c.items() if k.split(".")[0] in ("LDS", "STS", "LDG", "STG", "ATOMG", "ATOMS", "LDC", "LDCU", "RED"))
            print("  loop 0x%04x..0x%04x: %d instr  ALU %d  FMAH(slots) %d  LSU %d  other %d" % (ins[lo][0], a, len(body), alu, fma, lsu, len(body) - alu - lsu - sum(v for k, v in c.items() if k.startswith("IMAD") or k == "VIADD")))
            print("   ", dict(c.most_common()))
