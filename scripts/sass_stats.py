#!/usr/bin/env python
"""Static SASS statistics of a kernel in libg2048.so: per-pipe instruction counts of the hot loop,
with forward-branch-guarded blocks listed separately (they are skipped on the common path)."""
import collections
import re
import subprocess
import sys

ALU = {"LOP3", "PRMT", "SHF", "IADD3", "VIADD", "ISETP", "SEL", "PLOP3", "LEA", "POPC", "FLO", "BREV", "VIMNMX", "IABS", "FMNMX", "FSETP", "MOV", "BMSK", "SGXT", "I2FP", "F2I"}
FMA = {"IMAD", "FADD", "FMUL", "FFMA", "HFMA2"}


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
    def op(ins):
        t = ins.split()
        o = t[1] if t[0].startswith("@") else t[0]
        return o.split(".")[0]
    # the hot loop = the backward branch that spans the most instructions without enclosing another backward
    # branch's whole loop plus code outside it (barrier spin loops are short; the epilogue is not in any loop)
    back = []
    for a, i in body:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)\s*$", i) if op(i) == "BRA" else None
        if m and int(m.group(1), 16) < a:
            back.append((a - int(m.group(1), 16), a, int(m.group(1), 16)))
    inner = [b for b in back if not any(o is not b and b[2] <= o[2] and o[1] <= b[1] and o[0] > 64 for o in back)]
    _, loop_end, tgt = max(inner if inner else back)
    loop = [(a, i) for a, i in body if tgt <= a <= loop_end]
    # forward branches inside loop define skippable blocks
    blocks = []
    for a, i in loop:
        if op(i) == "BRA" and a != loop_end:
            t = int(re.search(r"0x([0-9a-f]+)\s*$", i).group(1), 16)
            if t > a:
                blocks.append((a, t, i))
    skipped = set()
    for a, t, i in blocks:
        for b, _ in loop:
            if a < b < t:
                skipped.add(b)
    def count(ins_list):
        c = collections.Counter()
        for _, i in ins_list:
            o = op(i)
            c["ALU" if o in ALU else "FMA" if o in FMA else "OTHER:" + o] += 1
        return c
    common = [(a, i) for a, i in loop if a not in skipped]
    print("kernel", pat, "total", len(body), "loop", len(loop), "common-path", len(common))
    c = count(common)
    print("  common path: ALU %d  FMA %d  other %d  %s" % (c["ALU"], c["FMA"], len(common) - c["ALU"] - c["FMA"],
          {k: v for k, v in c.items() if k.startswith("OTHER")}))
    for a, t, i in blocks:
        blk = [(b, j) for b, j in loop if a < b < t]
        cb = count(blk)
        print("  block %04x-%04x (%s): %d instr, ALU %d FMA %d" % (a, t, i, len(blk), cb["ALU"], cb["FMA"]))
    ops = collections.Counter(op(i) for _, i in common)
    print("  ", ops.most_common(20))


if __name__ == "__main__":
    main()
