#!/usr/bin/env python3
"""Turns the captures of tools/prof_r02.sh (gpurun_out/r02_*.ncu-rep, r02_launches_bench.csv) into the
committed evidence: profiles/r02_ncu_summary.txt, profiles/r02_launches_bench_summary.txt and
profiles/traffic.json (DRAM bytes per launch of every kernel, read by bench.py's `rooflines`)."""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "dram__sectors_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum"]
MULT = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}
NAMES = [("map", "map"), ("reduce", "reduce"), ("scan", "scan"), ("compress_p50", "compress"), ("compress_p01", "compress_p01"), ("compress_p99", "compress_p99"),
         ("hist", "histogram"), ("hist_fold", "histogram_fold"), ("gather_dram", "gather_dram"), ("gather_l2", "gather_l2")]
traffic = {"_source": "profiles/r02_ncu_summary.txt (tools/prof_r02.sh: ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum of one launch at the size bench.py runs)"}
lines = []
for rep, key in NAMES:
    path = os.path.join(OUT, f"r02_{rep}.ncu-rep")
    if not os.path.exists(path):
        lines.append(f"== {rep}: no capture"); continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        lines.append(f"== {rep}: empty capture"); continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines.append(f"== {rep}: {d.get('Kernel Name', ('?', ''))[0]}")
    rd = wr = None
    for k in KEYS:
        if k in d:
            v, u = d[k]
            lines.append(f"   {k:86s} {v} {u}")
            if k == "dram__bytes_read.sum": rd = float(v.replace(",", "")) * MULT.get(u.strip().lower(), 1)
            if k == "dram__bytes_write.sum": wr = float(v.replace(",", "")) * MULT.get(u.strip().lower(), 1)
    if rd is not None and wr is not None:
        traffic[f"{key}_dram_bytes_per_launch"] = int(rd + wr)
        lines.append(f"   DRAM traffic per launch: {(rd + wr) / 1e9:.4f} GB")
open(os.path.join(ROOT, "profiles", "r02_ncu_summary.txt"), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=2)
# launch list: share of every kernel in the bench command
path = os.path.join(OUT, "r02_launches_bench.csv")
if os.path.exists(path):
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:]))) if start >= 0 else []
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v / 1e3
        name = r["Kernel Name"].split("(")[0][-70:]
        agg[name][0] += 1
        agg[name][1] += v_us
    total = sum(v[1] for v in agg.values()) or 1
    out = [f"# python bench.py --steps 3 --warmup 3 under ncu (gpu__time_duration.sum, cold-cache, serialised): {len(rows)} launches, {total / 1e3:.1f} ms of GPU time",
           "# share of GPU time | launches | total us | kernel"]
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{100 * us / total:6.2f} %  {n:5d}  {us:12.1f}  {name}")
    open(os.path.join(ROOT, "profiles", "r02_launches_bench_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(lines[:60]))
