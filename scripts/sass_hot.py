#!/usr/bin/env python
"""Static look at the native kernel's SASS (no GPU needed): prints every plate-appearance step
(a run of ten IMAD.WIDE = one Philox2x32-10 block, up to the next BSYNC) with its instruction count
split by issue pipe, for the first template instance that has ATOMS (COUNT_PA) and no fatigue loop.

    python scripts/sass_hot.py [object file] [--dump]
"""
import re
import subprocess
import sys

obj = next((a for a in sys.argv[1:] if a.endswith(".o")), "mlb_odds_engine_b200/build/native_kernel.o")
out = subprocess.run(["cuobjdump", "-sass", "-fun", "mlbsim_native_kernel", obj], capture_output=True, text=True).stdout
ins = []
for line in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append(m.group(2).strip())
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2")
LSU = ("LDS", "STS", "ATOMS", "LDG", "STG", "RED", "ATOMG", "LDC", "LDCU")
CTRL = ("BRA", "BSSY", "BSYNC", "WARPSYNC", "NOP", "EXIT", "BAR", "VOTE", "S2R", "CS2R", "SHFL", "REDUX", "POPC")


def pipe(op):
    op = op.split()[1] if op.startswith("@") else op.split()[0]
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if base in LSU:
        return "lsu"
    if base in CTRL:
        return "ctl"
    return "alu"


# PA steps: find runs containing 10 IMAD.WIDE.U32 between BSSY ... BSYNC
steps = []
i = 0
n = len(ins)
while i < n:
    if "IMAD.WIDE.U32" in ins[i]:
        # back up to the preceding BSSY
        j = i
        while j > 0 and "BSSY" not in ins[j] and i - j < 6:
            j -= 1
        k = i
        wides = 0
        while k < n and "BSYNC" not in ins[k]:
            wides += "IMAD.WIDE.U32" in ins[k]
            k += 1
        steps.append((j, k, wides))
        i = k + 1
    else:
        i += 1
print(f"{len(ins)} SASS instructions; {len(steps)} Philox-bearing segments")
steps = [t for t in steps if t[2] == 10]
for (j, k, wides) in steps:
    seg = ins[j:k + 1]
    c = {"alu": 0, "fma": 0, "lsu": 0, "ctl": 0}
    for s in seg:
        c[pipe(s)] += 1
    print(f"  [{j}:{k}] {len(seg)} instr, IMAD.WIDE {wides}: {c}")
if "--dump" in sys.argv and steps:
    cand = [t for t in steps if any("ATOMS" in x for x in ins[t[0]:t[1] + 1])]
    j, k, _ = min(cand, key=lambda t: t[1] - t[0])
    for s in ins[j:k + 1]:
        print("     ", pipe(s), s)

if "--range" in sys.argv:
    a, b = (int(x) for x in sys.argv[sys.argv.index("--range") + 1].split(":"))
    for idx in range(a, b):
        print(f"{idx:5d} {pipe(ins[idx]):4s} {ins[idx]}")
