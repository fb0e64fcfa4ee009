"""Summarise an `ncu --page source --csv` dump: where the stall samples and instructions go.

  ncu -i X.ncu-rep --page source --csv [--kernel-name regex:...] > src.csv
  python scripts/ncu_src_summary.py src.csv [top_n_lines] [section_index]

The dump holds one section per profiled launch ("Kernel Name" row, header row, one row per SASS
instruction); section_index picks one (default 0).
"""
import csv
import sys


def sections(path):
    out = []
    cur = None
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1] if len(r) > 1 else "", "hdr": None, "rows": []}
            out.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) >= len(cur["hdr"]) - 1 and r:
            cur["rows"].append(r)
    return out


def num(s):
    try:
        return int(s or 0)
    except ValueError:
        try:
            return int(float(s))
        except ValueError:
            return 0


def main():
    secs = sections(sys.argv[1])
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    s = secs[which]
    hdr, data = s["hdr"], s["rows"]
    ix = {h: i for i, h in enumerate(hdr)}
    print("kernel:", s["name"][:100], "(%d sections in file)" % len(secs))
    tot_samples = sum(num(r[ix["# Samples"]]) for r in data)
    tot_inst = sum(num(r[ix["Instructions Executed"]]) for r in data)
    tot_thr = sum(num(r[ix["Thread Instructions Executed"]]) for r in data)
    print("instructions executed (warp): %d  (thread): %d   samples: %d" % (tot_inst, tot_thr, tot_samples))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(num(r[ix[h]]) for r in data) for h in stall_cols}
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    ops = {}
    for r in data:
        toks = r[ix["Source"]].split()
        op = toks[0] if toks else "?"
        if op.startswith("@") and len(toks) > 1:
            op = toks[1]
        o = ops.setdefault(op.split(".")[0], [0, 0])
        o[0] += num(r[ix["Instructions Executed"]])
        o[1] += num(r[ix["# Samples"]])
    print("%-12s %14s %7s %10s %7s" % ("opcode", "warp-instr", "%", "samples", "%"))
    for op, (n, sm) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
        print("%-12s %14d %6.2f%% %10d %6.2f%%" % (op, n, 100.0 * n / max(1, tot_inst), sm, 100.0 * sm / max(1, tot_samples)))
    if top_n:
        top = sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:top_n]
        for r in top:
            st = {h[6:]: num(r[ix[h]]) for h in stall_cols if num(r[ix[h]])}
            print(r[ix["Address"]][-5:], "%-70s" % r[ix["Source"]].strip()[:70], r[ix["# Samples"]], r[ix["Instructions Executed"]], st)


if __name__ == "__main__":
    main()
