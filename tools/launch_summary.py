"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: launches, total us, share.

    python tools/launch_summary.py gpurun_out/launches.csv ["header line for the summary"]
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    hdr, tot, cnt, order = None, collections.Counter(), collections.Counter(), []
    for r in csv.reader(open(path, errors="replace")):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d["Metric Name"] != "gpu__time_duration.sum":
                continue
            name = re.sub(r"\(.*", "", d["Kernel Name"])[:72]
            us = float(d["Metric Value"].replace(",", "")) / (1e3 if d["Metric Unit"] in ("ns", "nsecond") else 1.0)
            tot[name] += us
            cnt[name] += 1
            order.append((name, us))
    total = sum(tot.values())
    if len(sys.argv) > 2:
        print(sys.argv[2])
    print(f"{len(order)} launches, {total / 1e3:.3f} ms serialised (cold-cache per-launch times)")
    print(f"{'kernel':74s} {'n':>4s} {'us':>9s} {'share':>6s}")
    for name, us in tot.most_common():
        print(f"{name:74s} {cnt[name]:4d} {us:9.1f} {100 * us / total:5.1f}%")


if __name__ == "__main__":
    main()
