"""Turns the ncu artefacts brought back in gpurun_out/ into the small summaries committed here.
usage: python profiles/summarize.py gpurun_out/r01_launches.csv gpurun_out/r01_full.ncu-rep r01"""
import collections
import csv
import subprocess
import sys

launch_csv, rep, tag = sys.argv[1:4]
# ---- launch list: per-kernel time share (cold-cache, serialised: compare shares, not absolutes)
rows = []
with open(launch_csv) as f:
    for r in csv.reader(l for l in f if not l.startswith("==")):
        rows.append(r)
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", "")) / (1000.0 if r[hdr.index("Metric Unit")] in ("ns", "nsecond") else 1.0)
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_ncu_launches_summary.md", "w") as out:
    out.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over bench.py --batch 128\n\n")
    out.write("Per-launch times under ncu are cold-cache and serialised; the SHARE is what must agree with bench.py's CUDA-event shares.\n\n")
    out.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| {k} | {n} | {us:.1f} | {us / tot:.3f} |\n")
# ---- full capture: DRAM traffic and the main limiter per kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
idx = {n: i for i, n in enumerate(h)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
with open(f"profiles/{tag}_ncu_full_summary.md", "w") as out:
    out.write(f"# ncu --set full ({tag}), bench.py --batch 128, one launch per kernel\n\n| kernel | " + " | ".join(w.split(".")[0].replace("__", " ") for w in want) + " |\n|---|" + "---|" * len(want) + "\n")
    seen = set()
    for r in rr[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        if name in seen:
            continue
        seen.add(name)
        out.write(f"| {name} | " + " | ".join(f"{r[idx[w]]} {units[idx[w]]}" if w in idx else "-" for w in want) + " |\n")
print(open(f"profiles/{tag}_ncu_full_summary.md").read())
print(open(f"profiles/{tag}_ncu_launches_summary.md").read())
