#!/usr/bin/env python3
"""registers / stack / spills of the kernels whose names contain any of the given substrings, from
katome_b200/lib/ptxas.log (written by the Makefile: nvcc -Xptxas -v)"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(os.path.join(ROOT, "katome_b200", "lib", "ptxas.log")).read()
ents = re.findall(r"Compiling entry function '(\w+)' for 'sm_100a'\n.*?Function properties for \1\n\s+(\d+) bytes stack frame, "
                  r"(\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", log, re.S)
names = subprocess.run(["cu++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.splitlines() if ents else []
pats = sys.argv[1:]
for n, e in sorted(zip(names, ents)):
    if not pats or any(p in n for p in pats):
        print(f"{e[4]:>4} regs  stack {e[1]:>4}  spill st {e[2]:>4} ld {e[3]:>4}  {n[:140]}")
