"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (time share per kernel)."""
import collections
import csv
import sys


def main(path, skip_before=None):
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    # keep the launches of the LAST optimizer step only when a marker kernel is given
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in rows:
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        name = row["Kernel Name"]
        name = name.replace("dgfdn::<unnamed>::", "")
        agg[name[:90]][0] += 1
        agg[name[:90]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised; compare shares)")
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v / tot * 100:6.2f}%  {c:5d}x  {v / 1e3:10.1f} us  {n}")


if __name__ == "__main__":
    main(sys.argv[1])
