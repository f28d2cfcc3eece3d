"""Per-kernel DRAM traffic, warp-instruction count and issue utilisation of one captured frame (an `ncu --set full` report) as JSON:
what bench.py reads for `roofline.traffic` / `roofline.issue` (committed under profiles/, labelled as such in the bench line).
usage: python tools/ncu_traffic.py <report.ncu-rep> <sequences in the captured launches> <out.json> [note]"""
import csv
import json
import subprocess
import sys

rep, seqs, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale_units=True):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    if scale_units:
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    return v


kern = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip()
    if name.startswith("k_knn_cell_assoc<"):  # the unseeded and the seeded instantiation are the two launches of one frame
        name = "k_knn_cell_assoc"
    e = kern.setdefault(name, dict(dram_bytes_per_launch=[], warp_instructions_per_launch=[], issue_active_pct=[], us_under_ncu=[], threads_per_warp_instruction=[], sequences=seqs))
    e["dram_bytes_per_launch"].append(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
    e["warp_instructions_per_launch"].append(val(r, "smsp__inst_executed.sum", False))
    e["issue_active_pct"].append(val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False))
    e["threads_per_warp_instruction"].append(val(r, "smsp__thread_inst_executed_per_inst_executed.ratio", False))
    e["us_under_ncu"].append(val(r, "gpu__time_duration.sum"))
json.dump(dict(source=f"{rep.split('/')[-1]} (ncu --set full --clock-control none, one steady-state frame, launches in order) {note}", kernels=kern), open(out, "w"), indent=1)
for k, e in kern.items():
    print(f"{k:32s} launches {len(e['us_under_ncu'])}  us {sum(e['us_under_ncu']):8.1f}  dram MB {sum(e['dram_bytes_per_launch']) / 1e6:8.2f}  Minst {sum(e['warp_instructions_per_launch']) / 1e6:7.2f}")
