#!/usr/bin/env python3
"""Instruction mix of one kernel of libcsim_b200.so, per code window (static SASS, no GPU needed).

usage: tools/sass_mix.py <mangled-kernel-name> [window]
Prints, for consecutive windows of `window` instructions, the opcode histogram; FP64-dense windows are
the branch-free fast path of k_step_tb.  Used to check what ptxas made of the hot loop before spending
GPU time (profiles/*_tuning.md cite its output).
"""
import collections
import re
import subprocess
import sys

lib = __file__.rsplit("/", 2)[0] + "/climate-sim-mpi-cpp_b200/libcsim_b200.so"
fun = sys.argv[1]
win = int(sys.argv[2]) if len(sys.argv) > 2 else 250
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
ops = []
for line in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        t = m.group(2).split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops.append((int(m.group(1), 16), op.split(".")[0], m.group(2)))
print(len(ops), "instructions")
tot = collections.Counter(o for _, o, _ in ops)
print("total", dict(tot.most_common(12)))
for i in range(0, len(ops), win):
    c = collections.Counter(o for _, o, _ in ops[i:i + win])
    fp = c["DADD"] + c["DMUL"] + c["DFMA"]
    print(f"{ops[i][0]:#07x} fp64={fp:3d}", dict(c.most_common(8)))
if len(sys.argv) > 3:
    lo, hi = [int(x, 0) for x in sys.argv[3].split(":")]
    for a, _, txt in ops:
        if lo <= a < hi:
            print(f"{a:#07x}  {txt}")
