#!/usr/bin/env python
"""ptxas -v of one .cu as a table: kernel (demangled template arguments), registers, stack, spill stores / loads.
usage: python scripts/ptxas_summary.py trace.cu [extra nvcc flags]"""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(root, "6dof-pose-estimation-and-defect-projection_b200", "csrc", sys.argv[1])
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
       "-Xptxas", "-v", "-c", src, "-o", "/tmp/ptxas_summary.o"] + sys.argv[2:]
err = subprocess.run(cmd, capture_output=True, text=True).stderr
names = re.findall(r"Compiling entry function '(\S+)'", err)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
blocks = err.split("Compiling entry function ")[1:]
for d, b in zip(dem, blocks):
    regs = re.search(r"Used (\d+) registers", b)
    st = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    short = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "").replace("void ", ""))
    print(f"{short:60s} regs {regs.group(1):>3s}  stack {st.group(1):>4s}  spill st/ld {st.group(2):>4s}/{st.group(3):>4s}")
