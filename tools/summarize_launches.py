#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: the kernels of
ONE training step (between two consecutive embed_pe_fwd launches), grouped by kernel name.

    python tools/summarize_launches.py gpurun_out/launches.csv [title] > profiles/rNN_step_launches.txt
"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = next(r for r in rows if "Kernel Name" in r)
    data = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r is not hdr and r[0] != "ID"]
    starts = [i for i, d in enumerate(data) if "embed_pe_fwd" in d["Kernel Name"]]
    if len(starts) < 3:
        raise SystemExit("need at least three training steps in the capture")
    # the third and fourth occurrences bracket a full device-resident warm-up step (bench.py runs at least
    # four of them before anything else launches this kernel; later occurrences may belong to the e2e
    # loop, the roofline measurement or the baseline models)
    a, b = (starts[2], starts[3]) if len(starts) >= 4 else (starts[-3], starts[-2])
    step = data[a:b]
    agg = OrderedDict()
    for d in step:
        name = d["Kernel Name"].replace("void ", "").replace("etpgt::<unnamed>::", "")
        name = name.split("(")[0][:60]
        t = float(d["Metric Value"].replace(",", "")) / 1e3
        n, tot = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, tot + t)
    total = sum(t for _, t in agg.values())
    title = sys.argv[2] if len(sys.argv) > 2 else "one training step"
    print(f"{title}: {len(step)} launches, {total:.1f} us summed under ncu (cold-cache, serialised; compare SHARES)")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.1f} us {100 * t / total:5.1f}% x{n:3d}  {name}")


if __name__ == "__main__":
    main()
