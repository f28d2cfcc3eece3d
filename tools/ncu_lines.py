"""Hot source lines of one kernel in an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py <rep> <kernel-id e.g. ::k_voxel_cluster:1> [top]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; fname = ""
agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        try:
            smp = int(r[6]); ie = int(r[7])
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[4:]))
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
        agg.append((smp, ie, fname, r[0], r[1].strip()[:100], stalls))
tot = sum(a[0] for a in agg) or 1; toti = sum(a[1] for a in agg) or 1
print(f"# {kid}: {tot} samples, {toti} warp instructions")
for a in sorted(agg, key=lambda a: -a[0])[:top]:
    st = " ".join(f"{k}:{v}" for k, v in sorted(a[5].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*a[0]/tot:5.1f}% smp {100*a[1]/toti:5.1f}% inst  {a[2]}:{a[3]:>4}  {a[4]}\n        {st}")
