#!/usr/bin/env python
"""Attribute the SASS of one kernel of libschw_b200.so to source lines (needs -lineinfo; no GPU).

    python tools/sass_by_line.py 'persistent_raysINS_9RenderJobELb1ELb1' [--file schw_device.cuh --lines 250:360] [--dump]

Prints, per source line, how many instructions (and how many FP64-pipe instructions) the kernel holds for it,
summed over inlined copies; --dump lists the instructions themselves.  Used to check the per-step instruction
budget of the step loop before spending GPU time (B200_PROFILING.md: cuobjdump / nvdisasm first).
"""
import argparse
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200", "lib", "libschw_b200.so")
FP64 = ("DFMA", "DMUL", "DADD", "DSETP")

ap = argparse.ArgumentParser()
ap.add_argument("kernel", help="substring of the mangled kernel name")
ap.add_argument("--lib", default=LIB)
ap.add_argument("--file", default="schw_device.cuh")
ap.add_argument("--lines", default="1:100000")
ap.add_argument("--dump", action="store_true")
a = ap.parse_args()
lo, hi = [int(v) for v in a.lines.split(":")]

with tempfile.TemporaryDirectory() as d:
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(a.lib)], cwd=d, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, capture_output=True, text=True).stdout

inside = False
cur = None
cnt = collections.Counter(); fp = collections.Counter(); ops = collections.Counter()
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        inside = a.kernel in m.group(1)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.search(r"^\s+/\*([0-9a-f]+)\*/\s+((?:@!?U?P\d\s+)?)([A-Z0-9_.]+)(.*?);", line)
    if m and cur and cur[0] == a.file and lo <= cur[1] <= hi:
        op = m.group(3).split(".")[0]
        cnt[cur[1]] += 1
        ops[op] += 1
        if op in FP64:
            fp[cur[1]] += 1
        if a.dump:
            print(f"{cur[1]:5d} {m.group(1)} {m.group(2)}{m.group(3)}{m.group(4)}")
print("line  instr  fp64")
for k in sorted(cnt):
    print(f"{k:5d} {cnt[k]:6d} {fp[k]:5d}")
print("total", sum(cnt.values()), "fp64", sum(fp.values()))
print(dict(ops.most_common()))
