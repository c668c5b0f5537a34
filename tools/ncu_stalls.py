#!/usr/bin/env python
"""Side-by-side digest of `ncu --set full` captures of the product kernel:

    python tools/ncu_stalls.py a.ncu-rep [b.ncu-rep ...]

For the first kernel of each report: duration, FP64/DMMA pipe activity, issue-stall shares
(pc sampling), L2 hit rate, DRAM bytes, L1/shared-memory wavefront load and bank conflicts.
Written to compare the HBM-streamed default schedule with an L2-served (windowed) launch
(the L2-served stream is 13 % slower; this is the tool that has to say why).
"""
import csv
import io
import subprocess
import sys

SCALARS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "FP64/DMMA pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "shared ld wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "shared st wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "shared ld bank conflicts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "shared st bank conflicts"),
    ("launch__registers_per_thread", "registers / thread"),
]
STALL_PREFIX = "smsp__pcsamp_warps_issue_stalled_"


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, first = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, first)}


def main():
    reps = sys.argv[1:]
    if not reps:
        sys.exit(__doc__)
    data = [load(p) for p in reps]
    w = 46
    print(" " * w + "".join(f"{p.split('/')[-1][:24]:>26}" for p in reps))
    print(f"{'kernel':<{w}}" + "".join(f"{d.get('Kernel Name', ('?', ''))[0][:24]:>26}" for d in data))
    for key, label in SCALARS:
        cells = []
        for d in data:
            v, u = d.get(key, ("", ""))
            cells.append(f"{v} {u}".strip()[:24])
        print(f"{label:<{w}}" + "".join(f"{c:>26}" for c in cells))
    print("-- issue-stall shares (pc sampling), % of samples")
    shares = []
    for d in data:
        s = {}
        for k, (v, _) in d.items():
            if k.startswith(STALL_PREFIX) and not k.endswith("_not_issued"):
                try:
                    s[k[len(STALL_PREFIX):]] = float(v)
                except ValueError:
                    pass
        tot = sum(s.values()) or 1.0
        shares.append({k: 100.0 * v / tot for k, v in s.items()})
    names = sorted({k for s in shares for k in s}, key=lambda k: -max(s.get(k, 0.0) for s in shares))
    for k in names[:12]:
        print(f"  {k:<{w - 2}}" + "".join(f"{s.get(k, 0.0):>25.2f}%" for s in shares))


if __name__ == "__main__":
    main()
