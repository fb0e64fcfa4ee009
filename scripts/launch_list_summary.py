"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into profiles/rNN_launches.md.

  python scripts/launch_list_summary.py profiles/r01_launches.csv > profiles/r01_launches.md
"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
seq = []
for r in rows[1:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("rtc::", "")
    seq.append((int(r[ix["ID"]]), name, r[ix["Grid Size"]], r[ix["Block Size"]], float(r[ix["Metric Value"]])))
want = ["hoist_kernel", "trace_kernel", "shade_kernel", "count_kernel", "emit_kernel"]
frames, i = [], 0
while i + 4 < len(seq):
    if [s[1].split("<")[0] for s in seq[i:i + 5]] == want:
        frames.append(seq[i:i + 5]); i += 5
    else:
        i += 1
# frames of one kind only (same trace grid/template): the most common signature
sig = collections.Counter(tuple((k[1], k[2]) for k in f) for f in frames).most_common(1)[0][0]
main = [f for f in frames if tuple((k[1], k[2]) for k in f) == sig]
out = ["# Launch list: `python bench.py --steps 3 --warmup 3 --no-cpu-baseline` under ncu", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv` on a B200 (gpurun); raw list next to this file.",
       "Per-launch times are cold-cache and serialised (ncu flushes caches between launches): compare SHARES, not absolutes.", "",
       "## One frame of config3_4k_1024 (3841x2160, 1024 spheres + plane, RGB_PIXEL), every ray against every sphere: 5 launches", "",
       "Mean over %d frames of the capture." % len(main), "",
       "| kernel | grid | block | mean us | share of the frame |", "|---|---|---|---|---|"]
tot = sum(sum(f[k][4] for f in main) / len(main) for k in range(5))
for k in range(5):
    m = sum(f[k][4] for f in main) / len(main)
    out.append("| %s | %s | %s | %.2f | %.1f %% |" % (main[0][k][1], main[0][k][2], main[0][k][3], m / 1e3, 100 * m / tot))
out.append("| **frame** | | | **%.2f** | |" % (tot / 1e3))
other_frames = [f for f in frames if f not in main]
if other_frames:
    out += ["", "## The same frame with per-tile sphere culling (`with_culling` leg of bench.py)", "",
            "| kernel | mean us |", "|---|---|"]
    sig2 = collections.Counter(tuple((k[1], k[2]) for k in f) for f in other_frames).most_common(1)[0][0]
    oth = [f for f in other_frames if tuple((k[1], k[2]) for k in f) == sig2]
    for k in range(5):
        out.append("| %s | %.2f |" % (oth[0][k][1], sum(f[k][4] for f in oth) / len(oth) / 1e3))
in_frames = {s[0] for f in frames for s in f}
rest = collections.OrderedDict()
for s in seq:
    if s[0] not in in_frames:
        a = rest.setdefault((s[1][:64], s[2]), [0, 0.0]); a[0] += 1; a[1] += s[4]
out += ["", "## Everything else in the capture", "", "| kernel | grid | launches | mean us | what |", "|---|---|---|---|---|"]
for (k, g), (n, t) in rest.items():
    what = ("FP32 peak microbenchmark (rtc_fp32_peak, after the timed region)" if "peak" in k else
            "encoder stress, config 5 (7681x4320 random RGB)" if k.startswith(("count_kernel", "emit_kernel")) else
            "torch: L2 flush memset / test data (untimed)" if k.startswith("at::") else "")
    out.append("| %s | %s | %d | %.2f | %s |" % (k, g, n, t / n / 1e3, what))
print("\n".join(out))
