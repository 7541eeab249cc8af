#!/usr/bin/env python
"""Summarise an `ncu --set full` capture into the small JSON kept under profiles/ (development aid).

usage: ncu -i capture.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv profiles/rN_ncu_full.json

One entry per distinct kernel (first launch seen): duration, DRAM bytes read/written (the `traffic` of bench.py's
roofline object), DRAM / L2 / issue utilisation, occupancy, registers, instruction count, launch shape and the four
largest warp-stall reasons from the PC sampler.
"""
import collections
import csv
import json
import sys

WANT = {
    "dur_us": "gpu__time_duration.sum",
    "dram_rd_MB": "dram__bytes_read.sum",
    "dram_wr_MB": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "inst": "smsp__inst_executed.sum",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
}
BYTES = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
TIME_US = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}


def main(raw_csv: str, out_json: str) -> None:
    rows = list(csv.reader(open(raw_csv)))
    hdr, units = rows[0], rows[1]
    col = hdr.index
    stall = [i for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    out = collections.OrderedDict()
    for r in rows[2:]:
        name = r[col("Kernel Name")].split("(")[0].replace("void ", "").replace("kr::", "")
        if name in out:
            continue
        d = {}
        for key, metric in WANT.items():
            if metric not in hdr:
                continue
            x = float(r[col(metric)].replace(",", "") or 0)
            u = units[col(metric)]
            if key == "dur_us":
                x *= TIME_US.get(u, 1.0)
            if key in ("dram_rd_MB", "dram_wr_MB"):
                x *= BYTES.get(u, 1.0)
            d[key] = round(x, 3)
        tot = sum(float(r[i] or 0) for i in stall) or 1.0
        top = sorted(stall, key=lambda i: -float(r[i] or 0))[:4]
        d["top_stalls"] = [(hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", ""), round(100 * float(r[i] or 0) / tot)) for i in top]
        out[name] = d
    json.dump(out, open(out_json, "w"), indent=1)
    for k, v in out.items():
        print(k, v)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
