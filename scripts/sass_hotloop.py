#!/usr/bin/env python
"""Dump the hot loop of a step kernel from libg2048.so as annotated SASS: one line per instruction with the pipe it
issues to (A = ALU half-rate: LOP3/PRMT/SHF/ISETP/SEL..., F = FMA-heavy: IMAD, W = IMAD.WIDE/.HI, L = FMA-lite
capable: FADD, X = IADD3/VIADD (either integer pipe), M = memory, B = branch/barrier) followed by the counts
scripts/sass_stats.py prints.   python scripts/sass_hotloop.py [so] [kernel pattern] > profiles/rNN_step_sass_hotloop.txt"""
import re
import subprocess
import sys

ALU = {"LOP3", "PRMT", "SHF", "ISETP", "SEL", "PLOP3", "LEA", "POPC", "FLO", "BREV", "VIMNMX", "IABS", "FMNMX", "FSETP", "MOV", "BMSK", "SGXT", "I2FP", "F2I"}
FLEX = {"IADD3", "VIADD"}
MEM = {"LDG", "STG", "LDS", "STS", "LDC", "LDCU", "CCTL", "ATOMG", "RED"}


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else "gym-2048_b200/libg2048.so"
    pat = sys.argv[2] if len(sys.argv) > 2 else "step_kernelILj0ELb0ELi0E"
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    name, body = None, []
    for l in txt.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            name = m.group(1)
        elif name and pat in name:
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
            if m:
                body.append((int(m.group(1), 16), m.group(2).strip()))

    def opname(ins):
        t = ins.split()
        return t[1] if t[0].startswith("@") else t[0]
    back = []
    for a, i in body:
        o = opname(i)
        m = re.search(r"0x([0-9a-f]+)\s*$", i) if o.startswith("BRA") else None
        if m and int(m.group(1), 16) < a:
            back.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
    # the hot loop = the widest backward branch that does not enclose another loop plus code outside it (the
    # mbarrier retry path of the prologue jumps back over the whole kernel) — the rule scripts/sass_stats.py uses
    inner = [b for b in back if not any(o is not b and b[1] <= o[1] and o[2] <= b[2] and o[0] > 64 for o in back)]
    _, lo, hi = max(inner if inner else back)
    print("# %s: hot loop 0x%04x..0x%04x of %s" % (pat, lo, hi, so))
    print("# pipe tags: A ALU (half rate)  F FMA-heavy (IMAD)  W IMAD.WIDE/.HI  L FADD  X IADD3/VIADD  M memory  B branch/barrier  U uniform datapath")
    counts = {}
    for a, i in body:
        if not lo <= a <= hi:
            continue
        o = opname(i)
        b = o.split(".")[0]
        if b in ALU:
            tag = "A"
        elif b in FLEX:
            tag = "X"
        elif b == "IMAD":
            tag = "W" if (".WIDE" in o or ".HI" in o) else "F"
        elif b in ("FADD", "FMUL", "FFMA"):
            tag = "L"
        elif b in MEM:
            tag = "M"
        elif b.startswith("U") and b not in ("UMOV",):
            tag = "U"
        else:
            tag = "B"
        counts[tag] = counts.get(tag, 0) + 1
        print("%s  /*%04x*/  %s" % (tag, a, i))
    print("# loop instructions by tag (all blocks, skipped ones included): " + "  ".join("%s %d" % kv for kv in sorted(counts.items())))
    stats = subprocess.run([sys.executable, __file__.replace("sass_hotloop.py", "sass_stats.py"), so, pat], capture_output=True, text=True).stdout
    for l in stats.splitlines():
        print("# " + l)


if __name__ == "__main__":
    main()
