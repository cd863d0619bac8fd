"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys


def main(path, out=None):
    lines = open(path).readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in csv.DictReader(lines[start:]):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except Exception:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    rows = ["%-44s %7s %12s %10s %7s" % ("kernel", "launches", "total_us", "avg_us", "share")]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        rows.append("%-44s %7d %12.1f %10.2f %6.1f%%" % (k[:44], c, t, t / c, 100 * t / tot))
    rows.append("%-44s %7d %12.1f" % ("TOTAL", sum(c for c, _ in agg.values()), tot))
    text = "\n".join(rows)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
