#!/usr/bin/env python3
"""Reads an `ncu --page raw --csv` dump of the tconv_* launches of one bench.py step (tools/prof_tconv.sh) and
(1) writes the judged columns to profiles/<tag>_tconv_ncu_full.csv, (2) appends the DRAM traffic of the layer-0
launches (tconv_fwd + tconv_bwd_dst + tconv_bwd_src: dram__bytes_read.sum + dram__bytes_write.sum) keyed by the
batch shape to profiles/tconv_traffic.json, which bench.py reports as roofline.traffic.

    python tools/tconv_traffic.py raw.csv bench_log tag      # on the GPU box: writes gpurun_out/<tag>_tconv_ncu_full.csv
                                                             # and gpurun_out/<tag>_tconv_traffic_entry.json
    python tools/tconv_traffic.py --merge tag                # here: copies both into profiles/ (tconv_traffic.json)
    python tools/tconv_traffic.py --summary raw.csv out.csv  # the judged columns of any raw ncu dump
"""
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def merge(tag):
    out = ROOT / "gpurun_out"
    entry = json.loads((out / f"{tag}_tconv_traffic_entry.json").read_text())
    (ROOT / "profiles" / f"{tag}_tconv_ncu_full.csv").write_text((out / f"{tag}_tconv_ncu_full.csv").read_text())
    tfile = ROOT / "profiles" / "tconv_traffic.json"
    entries = json.loads(tfile.read_text()) if tfile.exists() else []
    entries = entries if isinstance(entries, list) else [entries]
    entries = [e for e in entries if e.get("capture") != entry["capture"]] + [entry]
    tfile.write_text(json.dumps(entries, indent=1) + "\n")
    print("merged", entry)


def summary(raw, out_csv):
    """The judged columns of any `ncu --page raw --csv` dump."""
    rows = [r for r in csv.reader(open(raw)) if r]
    hdr = next(r for r in rows if "Kernel Name" in r)
    units = rows[rows.index(hdr) + 1]
    data = [dict(zip(hdr, r)) for r in rows[rows.index(hdr) + 2:] if len(r) == len(hdr)]
    unit = dict(zip(hdr, units))
    keep = [k for k in KEEP if k in hdr]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([unit.get(k, "") for k in keep])
        for d in data:
            w.writerow([d[k] for k in keep])
    print(f"{len(data)} launches -> {out_csv}")


def main():
    if sys.argv[1] == "--merge":
        return merge(sys.argv[2])
    if sys.argv[1] == "--summary":
        return summary(sys.argv[2], sys.argv[3])
    raw, log, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(open(raw)) if r]
    hdr = next(r for r in rows if "Kernel Name" in r)
    units = rows[rows.index(hdr) + 1]
    data = [dict(zip(hdr, r)) for r in rows[rows.index(hdr) + 2:] if len(r) == len(hdr)]
    unit = dict(zip(hdr, units))
    keep = [k for k in KEEP if k in hdr]
    out_csv = ROOT / "gpurun_out" / f"{tag}_tconv_ncu_full.csv"
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([unit.get(k, "") for k in keep])
        for d in data:
            w.writerow([d[k] for k in keep])

    def dram(d):
        total = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(d[k].replace(",", "")) * SCALE.get(unit[k], 1.0)
        return total

    def first(pattern, which):
        hits = [d for d in data if pattern in d["Kernel Name"]]
        return hits[which] if hits else None

    # step order: fwd L0, fwd L1, bwd_dst L1, bwd_src L1, bwd_dst L0, bwd_src L0
    fwd, dst, src = first("tconv_fwd", 0), first("tconv_bwd_dst", -1), first("tconv_bwd_src", -1)
    if not (fwd and dst and src):
        raise SystemExit("tconv_traffic: the capture does not hold one tconv_fwd / bwd_dst / bwd_src launch each")
    shape = None
    for line in open(log):
        m = re.search(r'"batch_shape": \{"nodes": (\d+), "edges": (\d+)\}', line)
        if m:
            shape = (int(m.group(1)), int(m.group(2)))
    if shape is None:
        raise SystemExit("tconv_traffic: the bench line (batch_shape) is missing from the log")
    entry = {"nodes": shape[0], "edges": shape[1], "dram_bytes_fwd": dram(fwd), "dram_bytes_bwd_dst": dram(dst),
             "dram_bytes_bwd_src": dram(src), "dram_bytes_fwd_bwd": dram(fwd) + dram(dst) + dram(src),
             "capture": f"profiles/{tag}_tconv_ncu_full.csv (ncu --set full --clock-control none, one launch each, layer 0, "
                        "bench default batch)"}
    (ROOT / "gpurun_out" / f"{tag}_tconv_traffic_entry.json").write_text(json.dumps(entry, indent=1) + "\n")
    print(json.dumps(entry))


if __name__ == "__main__":
    main()
