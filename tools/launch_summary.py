"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) per kernel name:
count, mean duration and share of the summed GPU time.  python tools/launch_summary.py X.csv "header line" > X_summary.txt"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = defaultdict(list)
    for r in rows[1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1.0)
        acc[r[kn]].append(v)
    total = sum(sum(v) for v in acc.values())
    if len(sys.argv) > 2:
        print(sys.argv[2])
    print("launch times are cold-cache and serialised: compare shares\n")
    for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:92]:<92} n={len(v):4d} mean={sum(v) / len(v):9.1f} us share={100.0 * sum(v) / total:5.1f}%")


if __name__ == "__main__":
    main()
