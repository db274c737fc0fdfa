"""Instruction mix of every kernel of one object file (cuobjdump -sass): tensor-core / TMA / TMEM / reduction mnemonics and the plain math.
usage: sass_mix.py /tmp/transoar_b200_build/conv3d_gen_capi.o > profiles/...txt"""
import collections
import re
import subprocess
import sys

KEEP = ("UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "HMMA", "REDG", "RED.", "ATOM", "SYNCS", "LDG", "STG", "LDS", "STS", "SHFL", "FFMA", "FMUL", "MUFU", "LDC", "ELECT")
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
dem = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
name, counts, total = None, None, 0
def flush():
    if name:
        print(f"{dem(name)}   [{total} instructions]")
        print("    " + ", ".join(f"{v} {k}" for k, v in sorted(counts.items())))
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        name, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        total += 1
        op = m.group(1)
        if op.startswith(KEEP):
            counts[op] += 1
flush()
