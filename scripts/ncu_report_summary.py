"""Turn an .ncu-rep (ncu --set full --import-source on) into a small markdown summary for profiles/.

  python scripts/ncu_report_summary.py gpurun_out/prof.ncu-rep [kernel-regex] > profiles/rNN_x.md
"""
import csv
import io
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_src_summary import num, sections  # noqa: E402

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    kre = sys.argv[2] if len(sys.argv) > 2 else None
    extra = ["--kernel-name", "regex:" + kre] if kre else []
    raw = ncu(["-i", rep, "--page", "raw", "--csv"] + extra)
    rows = list(csv.reader(io.StringIO(raw)))
    rows = [r for r in rows if len(r) > 10]
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# ncu summary of `%s`%s\n" % (os.path.basename(rep), (" (kernels matching `%s`)" % kre) if kre else ""))
    print("Captured with `ncu --set full --clock-control none --import-source on` on a B200 via gpurun; per-launch "
          "values (ncu serialises launches and replays each ~40x: durations are cold-cache, compare shares).\n")
    names = [r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("rtc::", "") for r in data]
    print("| metric | unit | " + " | ".join("%d:%s" % (i, n[:28]) for i, n in enumerate(names)) + " |")
    print("|---|---|" + "---|" * len(data))
    for m in METRICS:
        if m not in ix:
            continue
        print("| %s | %s | %s |" % (m, units[ix[m]], " | ".join(r[ix[m]] for r in data)))
    src = ncu(["-i", rep, "--page", "source", "--csv"] + extra)
    tmp = "/tmp/_ncu_src_%d.csv" % os.getpid()
    open(tmp, "w").write(src)
    seen = set()
    for k, s in enumerate(sections(tmp)):
        key = s["name"].split("(")[0]
        if key in seen or not s["hdr"]:
            continue
        seen.add(key)
        h, d = s["hdr"], s["rows"]
        jx = {c: i for i, c in enumerate(h)}
        tot_i = sum(num(r[jx["Instructions Executed"]]) for r in d)
        tot_t = sum(num(r[jx["Thread Instructions Executed"]]) for r in d)
        tot_s = sum(num(r[jx["# Samples"]]) for r in d)
        print("\n## SASS mix: `%s`\n" % key.replace("void ", ""))
        print("warp instructions %d, thread instructions %d, stall samples %d\n" % (tot_i, tot_t, tot_s))
        stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        agg = sorted(((sum(num(r[jx[c]]) for r in d), c[6:]) for c in stall_cols), reverse=True)
        print("stall reasons (samples): " + ", ".join("%s %d" % (c, v) for v, c in agg if v) + "\n")
        ops = {}
        for r in d:
            toks = r[jx["Source"]].split()
            op = toks[0] if toks else "?"
            if op.startswith("@") and len(toks) > 1:
                op = toks[1]
            o = ops.setdefault(op.split(".")[0], [0, 0])
            o[0] += num(r[jx["Instructions Executed"]])
            o[1] += num(r[jx["# Samples"]])
        print("| opcode | warp instr | % | samples % |\n|---|---|---|---|")
        for op, (n, sm) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:16]:
            print("| %s | %d | %.1f | %.1f |" % (op, n, 100.0 * n / max(1, tot_i), 100.0 * sm / max(1, tot_s)))
    os.unlink(tmp)


if __name__ == "__main__":
    main()
