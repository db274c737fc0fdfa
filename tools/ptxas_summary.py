"""Condense the `-Xptxas -v` reports the csrc Makefile writes (one per translation unit, outside the tree) into one table:
kernel (demangled, template arguments kept), registers, spill bytes, static shared memory.   python tools/ptxas_summary.py [OBJDIR]"""
import glob
import os
import re
import subprocess
import sys


def main():
    objdir = sys.argv[1] if len(sys.argv) > 1 else "/tmp/transoar_b200_build"
    rows = []
    for path in sorted(glob.glob(os.path.join(objdir, "*_capi.ptxas.txt"))):
        unit = os.path.basename(path).replace("_capi.ptxas.txt", "")
        text = open(path).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:ptxas info\s+: Function properties[^\n]*\n)?"
                             r"(?:\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n)?"
                             r"ptxas info\s+: Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", text):
            rows.append((unit, m.group(1), int(m.group(5)), int(m.group(3) or 0), int(m.group(4) or 0), int(m.group(7) or 0), int(m.group(2) or 0)))
    names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
    print(f"# ptxas -v of every kernel of libmsda3d.so (sm_100a, -O3): {len(rows)} entry points; "
          f"{sum(1 for r in rows if r[3] or r[4])} with register spills")
    print("# unit | registers | spill st/ld bytes | static smem bytes | stack | kernel")
    for (unit, _, regs, sst, sld, smem, stack), name in zip(rows, names):
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*$", "", name)
        print(f"{unit:11s} {regs:4d}  {sst:4d}/{sld:<4d} {smem:6d} {stack:5d}  {name}")


if __name__ == "__main__":
    main()
