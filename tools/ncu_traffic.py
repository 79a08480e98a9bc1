"""Reads one `ncu --set full` capture of a bench.py step and writes, per kernel, the measured DRAM traffic and the headline
counters: profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`) and a CSV summary for profiles/.

  ncu -i capture.ncu-rep --page raw --csv > raw.csv
  python tools/ncu_traffic.py raw.csv --config 3 --summary profiles/ncu_full_r02_c3_summary.csv

Per kernel NAME the LAST launch in the capture is kept (the capture skips the warm-up, so every launch is a steady-state one)."""
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
]
UNIT_TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
UNIT_TO_US = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}


def short(name: str) -> str:
    name = name.split("(")[0].replace("<unnamed>::", "").replace("void ", "").strip()
    return name.split("<")[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv")
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--summary", default=None)
    args = ap.parse_args()
    rows = list(csv.reader(l for l in open(args.raw_csv) if l.startswith('"')))
    header, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(header)}
    name_col = col["Kernel Name"]
    kernels = {}
    for r in body:
        kernels[short(r[name_col])] = r   # last launch wins

    def value(r, metric, table):
        if metric not in col or r[col[metric]] in ("", "n/a"):
            return None
        v = float(r[col[metric]].replace(",", ""))
        return v * table.get(units[col[metric]], 1.0)

    traffic, summary = {}, []
    for k, r in kernels.items():
        rd, wr = value(r, "dram__bytes_read.sum", UNIT_TO_BYTES), value(r, "dram__bytes_write.sum", UNIT_TO_BYTES)
        if rd is not None and wr is not None:
            traffic[k] = int(rd + wr)
        line = {"kernel": k, "dram_bytes": traffic.get(k), "duration_us": value(r, "gpu__time_duration.sum", UNIT_TO_US)}
        for m in KEEP[3:]:
            line[m] = value(r, m, {})
        summary.append(line)
    # names bench.py reports rooflines under
    if "tess_count_kernel" in traffic:
        traffic["tess_count+scan+emit"] = traffic["tess_count_kernel"] + traffic.get("fill_segments_kernel", 0) + traffic.get("tess_emit_kernel", 0)
    if "raster_tiles_kernel" in traffic and "tile_prims_kernel" in traffic:
        traffic["raster_tiles_kernel+tile_prims_kernel"] = traffic["raster_tiles_kernel"] + traffic["tile_prims_kernel"]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = {}
    if os.path.exists(path):
        with open(path) as f:
            table = json.load(f)
    table[f"config{args.config}"] = traffic
    with open(path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    if args.summary:
        keys = list(summary[0].keys())
        with open(args.summary, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=keys)
            w.writeheader()
            w.writerows(summary)
    for line in summary:
        print(f"{line['kernel']:28s} {line['duration_us'] or 0:9.1f} us  dram {((line['dram_bytes'] or 0) / 1e6):9.2f} MB  issue {line.get('sm__issue_active.avg.pct_of_peak_sustained_elapsed')} %  barrier stall {line.get('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio')}")


if __name__ == "__main__":
    main()
