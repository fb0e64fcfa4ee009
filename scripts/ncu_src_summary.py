"""Summarise an `ncu --page source --csv` dump: where the stall samples and instructions go."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_samples = sum(int(r[ix["# Samples"]] or 0) for r in data)
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("instructions executed (warp): %d   samples: %d" % (tot_inst, tot_samples))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
# opcode histogram by instructions executed and samples
ops = {}
for r in data:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    o = ops.setdefault(op.split(".")[0], [0, 0])
    o[0] += int(r[ix["Instructions Executed"]] or 0)
    o[1] += int(r[ix["# Samples"]] or 0)
print("%-12s %14s %7s %10s %7s" % ("opcode", "warp-instr", "%", "samples", "%"))
for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
    print("%-12s %14d %6.2f%% %10d %6.2f%%" % (op, n, 100.0 * n / tot_inst, s, 100.0 * s / max(1, tot_samples)))
if len(sys.argv) > 2:
    top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[: int(sys.argv[2])]
    for r in top:
        st = {h[6:]: int(r[ix[h]] or 0) for h in stall_cols if int(r[ix[h]] or 0)}
        print(r[ix["Address"]][-5:], "%-70s" % r[ix["Source"]].strip()[:70], r[ix["# Samples"]], r[ix["Instructions Executed"]], st)
