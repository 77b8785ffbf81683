"""Turns the ncu artefacts brought back in gpurun_out/ into the small summaries committed here.
usage: python profiles/summarize.py LAUNCHES.csv FULL.ncu-rep TAG STREAMS_PER_LAUNCH [WORKLOAD_DESCRIPTION]
writes profiles/TAG_ncu_launches_summary.md, profiles/TAG_ncu_full_summary.md and merges the kernels it saw into
profiles/ncu_traffic.json (DRAM bytes per launch and per stream of every kernel, read by bench.py for `roofline.traffic`)."""
import collections
import csv
import json
import subprocess
import sys

launch_csv, rep, tag, streams = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
what = sys.argv[5] if len(sys.argv) > 5 else "scanRegistration + laserOdometry + laserMapping on the 1 M-point corridor map, forward trajectory"
cmd = sys.argv[6] if len(sys.argv) > 6 else f"bench.py --workload sr_lo_lm --legs device --batch {streams} --handles 1 --steps 3 --warmup 4"
bench_json = sys.argv[7] if len(sys.argv) > 7 else "profiles/r02_bench_default.json"


def base(name):
    return name.split("(")[0].replace("void ", "").replace("vb::", "").split("<")[0].replace("_occ6", "").replace("_occ5", "").replace("_occ8", "").replace("_occ3", "")


# ---- launch list: per-kernel time share (cold-cache, serialised: compare shares, not absolutes)
rows = []
with open(launch_csv) as f:
    for r in csv.reader(l for l in f if not l.startswith("==")):
        rows.append(r)
hdr = rows[0]
ki, vi, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    us = float(r[vi].replace(",", ""))
    us = us / 1000.0 if r[ui] in ("ns", "nsecond") else us * (1000.0 if r[ui] in ("ms", "msecond") else 1.0)
    a = agg[base(r[ki])]
    a[0] += 1
    a[1] += us
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_ncu_launches_summary.md", "w") as out:
    out.write(f"# ncu launch list ({tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
              f"`{cmd}` ({what}).\n"
              "Per-launch times under ncu are cold-cache and serialised; the SHARE is what must agree with the CUDA-event shares in\n"
              f"`{bench_json}` (`kernels.*.share`).\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"| {k} | {n} | {us:.1f} | {us / tot:.3f} |\n")

# ---- full capture: DRAM traffic and the main limiter per kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
idx = {n: i for i, n in enumerate(h)}


def val(r, name):
    try:
        v = float(r[idx[name]].replace(",", ""))
    except (KeyError, ValueError):
        return None
    u = units[idx[name]]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0) if "byte" in u else v


cols = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "dram read MB"), ("dram__bytes_write.sum", "dram write MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "warp instr"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
traffic = {}
with open(f"profiles/{tag}_ncu_full_summary.md", "w") as out:
    out.write(f"# ncu --set full ({tag})\n\nOne scan of `{cmd}` ({what}; a steady-state scan); one row per launch.\n"
              "DRAM bytes are per launch (all streams of the launch).\n\n| kernel | "
              + " | ".join(c[1] for c in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
    for r in rr[2:]:
        name = base(r[idx["Kernel Name"]])
        cells = []
        for m, label in cols:
            v = val(r, m)
            if v is None:
                cells.append("-")
            elif "MB" in label:
                cells.append(f"{v / 1e6:.1f}")
            elif label == "time us":
                u = units[idx[m]]
                cells.append(f"{v / 1000.0 if u.startswith('n') else v:.1f}")
            elif label in ("regs", "warp instr"):
                cells.append(f"{int(v)}")
            else:
                cells.append(f"{v:.1f}")
        out.write(f"| {name} | " + " | ".join(cells) + " |\n")
        rd, wr = val(r, "dram__bytes_read.sum") or 0.0, val(r, "dram__bytes_write.sum") or 0.0
        t = traffic.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "issue": 0.0, "time_us": 0.0})
        t["launches"] += 1
        t["dram_bytes"] += rd + wr
        t["issue"] += val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0.0
        tu = val(r, "gpu__time_duration.sum") or 0.0
        t["time_us"] += tu / 1000.0 if units[idx["gpu__time_duration.sum"]].startswith("n") else tu
for k, t in traffic.items():
    t["dram_bytes_per_launch"] = t["dram_bytes"] / t["launches"]
    t["dram_bytes_per_launch_per_stream"] = t["dram_bytes_per_launch"] / streams
    t["issue_active_pct"] = t.pop("issue") / t["launches"]
    t["ncu_time_us_per_launch"] = t.pop("time_us") / t["launches"]
    t["streams_per_launch"] = streams
    del t["dram_bytes"]
for t in traffic.values():
    t["source"] = f"profiles/{tag}_ncu_full_summary.md"
try:
    merged = json.load(open("profiles/ncu_traffic.json"))
except Exception:
    merged = {"kernels": {}}
merged.setdefault("kernels", {}).update(traffic)
merged["source"] = "profiles/*_ncu_full_summary.md (per kernel: `source`)"
merged.pop("streams_per_launch", None)
json.dump(merged, open("profiles/ncu_traffic.json", "w"), indent=1)
print(open(f"profiles/{tag}_ncu_full_summary.md").read())
print(open(f"profiles/{tag}_ncu_launches_summary.md").read())
