#!/usr/bin/env python
"""ncu launch-list CSV (`ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file X.csv CMD`)
-> the text summary committed under profiles/: per-kernel totals and shares, then the codec's own
launches in order and their shares of the codec's GPU time.

    python profiles/launches.py gpurun_out/launches_r1d.csv "python bench.py --steps 2 ..." > profiles/r01d_bench_launches.txt
"""
import collections
import csv
import sys


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    i_grid, i_block = hdr.index("Grid Size"), hdr.index("Block Size")
    launches = []
    for r in rows[1:]:
        if len(r) <= i_val:
            continue
        v = float(r[i_val].replace(",", ""))
        u = r[i_unit]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        launches.append((r[i_name], us, r[i_grid], r[i_block]))
    total = sum(l[1] for l in launches)
    print("# ncu --metrics gpu__time_duration.sum --clock-control none %s" % cmd)
    print("# (B200; every launch of the process incl. synthetic-frame generation and torch checks; cold-cache, serialised)")
    print("# %d launches, %.1f us total" % (len(launches), total))
    print("# count   total_us   share  grid  block  kernel")
    agg = collections.OrderedDict()
    for name, us, g, b in launches:
        a = agg.setdefault(name, [0, 0.0, g, b])
        a[0] += 1
        a[1] += us
    for name, (n, us, g, b) in agg.items():
        print("%4d %12.1f  %5.1f%%  %s %s  %s" % (n, us, 100 * us / total, g, b, name[:110]))
    mine = [(n, us) for n, us, _, _ in launches if "dbde" in n]
    print("# the codec's own launches in order (us); one bench step = encode + decode_scan + decode:")
    for n, us in mine:
        print("   %10.1f  %s" % (us, n[:100]))
    tot = sum(us for _, us in mine) or 1.0
    per = collections.OrderedDict()
    for n, us in mine:
        key = n.split("(")[0].split("<")[0].split("dbde::")[-1].split()[-1]
        per[key] = per.get(key, 0.0) + us
    print("# shares inside the codec's launches: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in per.items()))


if __name__ == "__main__":
    main()
