#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, without a GPU) into the handful of numbers the roofline
needs: per launch duration, DRAM bytes read/written, DRAM/L2 throughput %, registers,
occupancy and the top stall reasons.

    python tools/ncu_summary.py gpurun_out/r01_scan.ncu-rep [more.ncu-rep ...] > profiles/....txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic",
    "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum",
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.strip().lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}
    return v * mult.get(u, 1)


def to_us(val, unit):
    v = float(val.replace(",", ""))
    u = unit.strip().lower()
    mult = {"ns": 1e-3, "us": 1, "usecond": 1, "msecond": 1e3, "ms": 1e3, "second": 1e6, "nsecond": 1e-3}
    return v * mult.get(u, 1)


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"{path}: no data")
            continue
        header, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(header)}
        print(f"== {path}")
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            print(f"-- {name[:110]}")
            dur = rd = wr = None
            for k in KEYS:
                if k in col:
                    v, u = r[col[k]], units[col[k]]
                    print(f"   {k:62s} {v} {u}")
                    if k == "gpu__time_duration.sum":
                        dur = to_us(v, u)
                    if k == "dram__bytes_read.sum":
                        rd = to_bytes(v, u)
                    if k == "dram__bytes_write.sum":
                        wr = to_bytes(v, u)
            if dur and rd is not None and wr is not None:
                print(f"   => DRAM traffic {rd + wr:.4g} B per launch, {(rd + wr) / dur / 1e3:.1f} GB/s over {dur:.1f} us (under ncu: cold, serialised)")
            stalls = []
            for h, i in col.items():
                if h.startswith("smsp__average_warp_latency_issue_stalled") or h.startswith("smsp__average_warps_issue_stalled"):
                    if h.endswith("_per_issue_active.ratio") or h.endswith(".ratio"):
                        try:
                            stalls.append((float(r[i].replace(",", "")), h))
                        except ValueError:
                            pass
            for v, h in sorted(stalls, reverse=True)[:6]:
                print(f"   stall {h.split('issue_stalled_')[-1].split('.')[0]:40s} {v:.2f}")


if __name__ == "__main__":
    main()
