"""profiles/rNN_ncu_traffic.json from .ncu-rep captures (ncu --set full): DRAM bytes per launch of each of our kernels --
what bench.py prints as roofline.traffic.  Nothing here is typed in by hand: every number is read from a capture and
carries the capture's file name.

  python scripts/ncu_traffic.py profiles/r02_ncu_traffic.json \
      gpurun_out/r2_prof_trace.ncu-rep:trace_kernel:config3_4k_1024 \
      gpurun_out/r2_prof_encode.ncu-rep:count_kernel:config5_encode_8k gpurun_out/r2_prof_encode.ncu-rep:emit_kernel:config5_encode_8k
"""
import csv
import io
import json
import os
import subprocess
import sys


def first_launch(rep, pattern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + pattern],
                         capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(raw)) if len(r) > 10]
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    r = data[0]

    def val(name):
        v = float(r[ix[name]].replace(",", ""))
        u = units[ix[name]].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u.split("/")[0], 1.0)
        return v * scale
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("rtc::", "")
    return name, val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("gpu__time_duration.sum")


def main():
    out_path, specs = sys.argv[1], sys.argv[2:]
    out = {}
    for spec in specs:
        rep, pattern, workload = spec.split(":")
        name, rd, wr, us = first_launch(rep, pattern)
        key = pattern if pattern in ("trace_kernel",) else name
        out[key] = {"kernel": name, "workload": workload, "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
                    "gpu_time_us": us, "source": "%s (ncu --set full --clock-control none, cold caches, first matching launch)" % os.path.basename(rep)}
    json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
