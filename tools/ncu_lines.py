"""Summarise an .ncu-rep: key raw metrics + instructions / stall samples per CUDA source line (needs -lineinfo)."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
per_unit = float(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] else None      # warp-samples in the launch: warps * samples
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:72s} {units[i]:14s} {[r[i] for r in data]}")
for i, h in enumerate(hdr):
    if "issue_stalled" in h and "per_issue_active" in h:
        v = [r[i] for r in data]
        if any(float(x) > 0.02 for x in v):
            print(f"{h:72s} {v}")
inst = float(data[0][hdr.index("smsp__inst_executed.sum")])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
agg = {}
tot = tots = 0
for r in csv.reader(src.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name"):
        continue
    if r[2] == "-" and r[0].strip().isdigit():
        try:
            e = int(r[7]); s = int(r[6])
        except ValueError:
            continue
        k = (cur, int(r[0]), r[1].strip()[:86])
        a = agg.get(k, (0, 0)); agg[k] = (a[0] + e, a[1] + s)
        tot += e; tots += s
scale = (inst / per_unit / tot) if per_unit else 100.0 / tot
print("instructions per %s: %.1f" % ("warp-sample" if per_unit else "100", inst / per_unit if per_unit else 100.0))
for (f, ln, s_), (e, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"{e * scale:8.1f} {100 * s / max(tots,1):5.1f}%smp  {f}:{ln}  {s_}")
