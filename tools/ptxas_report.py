"""Per-kernel register / spill report: python tools/ptxas_report.py <file.cu> [filter]"""
import re
import subprocess
import sys

src = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                      "-Xptxas", "-v", "-c", src, "-o", "/dev/null"], capture_output=True, text=True).stderr
cur = None
rows = []
for line in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        spill = (int(m.group(2)), int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        rows.append((cur, int(m.group(1)), spill))
        cur = None
for name, regs, spill in rows:
    if flt in name:
        print(f"{regs:4d} regs  spill st/ld {spill[0]:4d}/{spill[1]:4d}  {name}")
