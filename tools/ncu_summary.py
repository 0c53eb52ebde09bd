#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): headline counters, stall-reason totals,
the hottest SASS lines and per-opcode sample counts.   python tools/ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys
from collections import Counter

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sectors_op_read.sum", "sm__cycles_active.avg"]


def ncu(path, page):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if not l.startswith("==")]
    return list(csv.reader(io.StringIO("\n".join(lines))))


def main(path, ntop=18):
    raw = ncu(path, "raw")
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
        print("## kernel:", name[:110])
        for h, u, v in zip(hdr, units, row):
            if h in WANT:
                print(f"{h:75s} {v} {u}")
    src = ncu(path, "source")
    # the source page has a 'Kernel Name' row, then the header row
    hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h = src[hi]
    ix = {k: i for i, k in enumerate(h)}
    data = [r for r in src[hi + 1:] if len(r) == len(h)]
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    num = lambda r, k: int(float(r[ix[k]] or 0))
    total = sum(num(r, "# Samples") for r in data)
    print(f"## stall reasons ({total} samples)")
    for k, v in sorted(((k, sum(num(r, k) for r in data)) for k in stalls), key=lambda x: -x[1])[:8]:
        print(f"  {k:28s} {100.0 * v / max(total, 1):5.1f}%")
    print("## hottest SASS lines")
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:ntop]:
        st = {k: num(r, k) for k in stalls}
        b = max(st, key=st.get)
        print(f"  {num(r, '# Samples'):7d}  {r[ix['Source']].strip()[:60]:60s} {b} exec={r[ix['Instructions Executed']]}")
    ops, ex = Counter(), Counter()
    for r in data:
        t = r[ix["Source"]].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        ops[op] += num(r, "# Samples")
        ex[op] += num(r, "Instructions Executed")
    print("## per opcode: samples, warp-instructions executed")
    for k, v in ops.most_common(14):
        print(f"  {k:10s} {v:8d} {ex[k]:12d}")


if __name__ == "__main__":
    main(sys.argv[1])
