"""Per-source-line instruction / stall-sample totals from an ncu report captured with --import-source on.
usage: python profiles/srcview.py REPORT.ncu-rep KERNEL_REGEX [top_n]"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
cur_file, hdr = None, None
lines = {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or r[0] == "" or not r[0].isdigit():
        continue
    key = (cur_file, int(r[0]))
    try:
        inst, samp = int(r[i_inst] or 0), int(r[i_samp] or 0)
    except (ValueError, IndexError):
        continue
    a = lines.setdefault(key, [0, 0, r[1].strip()])
    a[0] += inst
    a[1] += samp
ti = sum(v[0] for v in lines.values()) or 1
ts = sum(v[1] for v in lines.values()) or 1
print(f"total warp instructions {ti}, stall samples {ts}")
print("by samples:")
for (f, ln), (inst, samp, src) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{samp / ts * 100:5.1f}% smp {inst / ti * 100:5.1f}% ins  {f}:{ln:<4d} {src[:110]}")
