"""Per-kernel table from an ncu --set full report: python tools/ncu_table.py gpurun_out/X.ncu-rep [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
cols = [("Kernel Name", "kernel", 34), ("gpu__time_duration.sum", "us", 8), ("launch__grid_size", "grid", 7), ("launch__block_size", "blk", 4),
        ("launch__registers_per_thread", "reg", 4), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 6),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 6), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 6),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 6), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%", 6),
        ("dram__bytes_read.sum", "rdMB", 7), ("dram__bytes_write.sum", "wrMB", 7),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/i", 6), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 6),
        ("smsp__inst_executed.sum", "Minst", 8), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 6),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stLSB", 6),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stBAR", 6),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stSSB", 6),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stMEMB", 6),
        ]
def conv(v, u, name):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if name == "us":
        x = x / 1000 if u in ("ns", "nsecond") else (x * 1000 if u in ("ms", "msecond") else x)
    if name in ("rdMB", "wrMB"):
        x = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1) * x
    if name == "Minst":
        x = x / 1e6
    return f"{x:.2f}" if abs(x) < 1000 else f"{x:.0f}"
lines = ["  ".join(n.rjust(w) if n != "kernel" else n.ljust(w) for _, n, w in cols)]
tot = 0.0
for row in r[2:]:
    out = []
    for key, n, w in cols:
        if key not in hdr:
            out.append("-".rjust(w)); continue
        i = hdr.index(key)
        v = row[i]
        if n == "kernel":
            out.append(v.split("(")[0][:w].ljust(w)); continue
        s = conv(v, units[i], n)
        if n == "us":
            tot += float(s)
        out.append(s.rjust(w))
    lines.append("  ".join(out))
lines.append(f"# total {tot:.1f} us over {len(r) - 2} launches")
txt = "\n".join(lines)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
