"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import re
import sys


def main(path, show_from=None):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg, seq = collections.OrderedDict(), []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1000, "us": v, "ms": v * 1000, "s": v * 1e6}.get(row["Metric Unit"], v)
        short = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")[:64]
        seq.append((short, v))
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    print(f"{len(seq)} launches, {tot / 1000:.2f} ms of device time")
    print(f"{'kernel':66s} {'n':>5s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:66s} {n:5d} {t:10.1f} {t / n:8.1f} {t / tot:6.3f}")
    if show_from is not None:
        for name, v in seq[show_from:show_from + 24]:
            print(f"    {name:62s} {v:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
