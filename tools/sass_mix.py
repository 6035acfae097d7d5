"""Instruction mix of the main loop of a kernel from cuobjdump -sass (static count between the loop head and the
largest loop that contains the gathers and the FMA work but not the final reduction).  Usage: python tools/sass_mix.py <obj-or-so> <substring of kernel name> [px per iteration]"""
import collections
import re
import subprocess
import sys

FMA = {"FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2", "IMAD", "HFMA2"}
XU = {"MUFU", "F2I", "I2F", "F2F", "I2FP"}
MEM = {"LDG", "STG", "LDS", "STS", "LDL", "STL", "TEX", "TLD", "UBLKCP", "SYNCS", "LDC", "LDCU", "ATOMG", "RED", "ATOMS"}


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    px = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)([^;]*);", f)
        addr = {int(a, 16): i for i, (a, _, _, _) in enumerate(ins)}
        best = None
        for i, (a, _, op, rest) in enumerate(ins):
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", rest)
                if m and int(m.group(1), 16) < int(a, 16):
                    span = int(a, 16) - int(m.group(1), 16)
                    lo_i = addr.get(int(m.group(1), 16), 0)
                    ops = [o.split(".")[0] for _, _, o, _ in ins[lo_i:i + 1]]
                    ok = ("TEX" in ops or "LDG" in ops) and "SHFL" not in ops and ops.count("FFMA") + 2 * ops.count("FFMA2") > 50
                    if ok and (best is None or span > best[0]):
                        best = (span, lo_i, i)
        if best is None:
            print(name, ": no loop"); continue
        _, lo, hi = best
        cnt = collections.Counter(op.split(".")[0] for _, _, op, _ in ins[lo:hi + 1])
        n = hi - lo + 1
        fma = sum(v * (2 if k in ("FFMA2", "FMUL2", "FADD2") else 1) for k, v in cnt.items() if k in FMA)
        fma_issue = sum(v for k, v in cnt.items() if k in FMA)
        xu = sum(v for k, v in cnt.items() if k in XU)
        mem = sum(v for k, v in cnt.items() if k in MEM)
        alu = n - fma_issue - xu - mem
        print("%s\n  loop %d instr = %.1f / px | fma-pipe issue %.1f (pipe cycles %.1f) | alu+ctl %.1f | xu %.1f | mem %.1f per px"
              % (name[:110], n, n / px, fma_issue / px, fma / px, alu / px, xu / px, mem / px))
        print("  " + ", ".join("%s %d" % kv for kv in cnt.most_common(40)))


if __name__ == "__main__":
    main()
