"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/.
usage: python tools/ncu_summary.py <tag> [note]"""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else ""
out = []
out.append(f"# ncu launch list  ({tag})  {note}")
out.append("# command: see tools/profile.sh; --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
with open(f"gpurun_out/{tag}_launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    agg[row["Kernel Name"]][0] += 1
    agg[row["Kernel Name"]][1] += v
tot = sum(v[1] for v in agg.values())
out.append(f"# {sum(v[0] for v in agg.values())} launches, {tot:.1f} us total")
out.append("share%  launches  avg_us  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{v[1] / tot * 100:5.1f}  {v[0]:5d}  {v[1] / v[0]:8.1f}  {k[:110]}")
open(f"profiles/{tag}_launches.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))

try:
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]
    txt = [f"# ncu --set full capture ({tag})  {note}", "# per launch; warp execution efficiency = smsp__thread_inst_executed_per_inst_executed.ratio / 32"]
    for row in r[2:]:
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                txt.append(f"{w} [{units[i]}] = {row[i]}")
        txt.append("")
    open(f"profiles/{tag}_full.txt", "w").write("\n".join(txt) + "\n")
    print("\n".join(txt[:30]))
except Exception as e:  # noqa: BLE001
    print("no full capture:", e)
