#!/usr/bin/env python
"""tools/ncu_summary.py -- condense an Nsight Compute report into the JSON kept under profiles/.

usage: ncu_summary.py REPORT.ncu-rep OUT.json [--traffic TRAFFIC.json]

Reads `ncu -i REPORT --page raw --csv` and keeps, per kernel launch, the metrics the roofline and the
issue-budget arguments in DESIGN.md are built on (duration, DRAM bytes read/written, DRAM and SM
throughput, executed warp instructions, issue-slot utilisation, registers, occupancy).  With
--traffic it also writes the per-launch dram__bytes_read.sum + dram__bytes_write.sum that bench.py
reports as roofline.traffic."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    m = re.search(r"(k_\w+)<\s*(?:nuk::)?(\w+)", name)
    return "%s<%s>" % (m.group(1), m.group(2)) if m else name.split("(")[0][:60]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_out = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    summary, traffic = {}, {}
    for r in data:
        name = short(r[col["Kernel Name"]])
        key, k = name, 1
        while key in summary:
            k += 1
            key = "%s#%d" % (name, k)
        ent = {"kernel": r[col["Kernel Name"]][:200], "unit_note": {}}
        for m in KEEP:
            if m in col and r[col[m]] != "":
                try:
                    ent[m] = float(r[col[m]].replace(",", ""))
                except ValueError:
                    continue
                ent["unit_note"][m] = units[col[m]]
        summary[key] = ent
        if "dram__bytes_read.sum" in ent and "dram__bytes_write.sum" in ent:
            b = (ent["dram__bytes_read.sum"] * TO_BYTES.get(ent["unit_note"]["dram__bytes_read.sum"], 1) +
                 ent["dram__bytes_write.sum"] * TO_BYTES.get(ent["unit_note"]["dram__bytes_write.sum"], 1))
            ent["dram_bytes_per_launch"] = int(b)
            traffic.setdefault(name, int(b))
            m = re.match(r"k_map_flat<(\w+)F>$", name)      # bench.py looks the C2 ufuncs up as k_map_flat<sin> ...
            if m:
                traffic.setdefault("k_map_flat<%s>" % m.group(1).lower(), int(b))
        if "smsp__inst_executed.sum" in ent:
            ent["thread_instructions_per_launch"] = ent["smsp__inst_executed.sum"] * 32
    json.dump(summary, open(out, "w"), indent=1)
    if traffic_out:
        traffic["_source"] = "ncu --set full --clock-control none (%s): dram__bytes_read.sum + dram__bytes_write.sum per launch" % rep
        json.dump(traffic, open(traffic_out, "w"), indent=1)
    for k, v in summary.items():
        print(k, v.get("gpu__time_duration.sum"), v.get("dram_bytes_per_launch"), v.get("smsp__inst_executed.sum"))


if __name__ == "__main__":
    main()
