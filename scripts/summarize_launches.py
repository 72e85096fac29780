"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (time share per kernel).

usage: summarize_launches.py launches.csv [marker]
With a marker (substring of a kernel name that runs once per step, e.g. td_fused_kernel) only the launches of ONE
step are kept -- from the second-to-last marker launch up to the last one -- so the shares can be compared with
bench.py's `roofline.share_of_step` (the whole list also holds the set-up, warm-up and capture launches)."""
import collections
import csv
import sys


def main(path, marker=None):
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    note = "all launches of the run"
    if marker:
        hits = [i for i, r in enumerate(rows) if marker in r.get("Kernel Name", "")]
        if len(hits) >= 2:
            rows = rows[hits[-2]:hits[-1]]
            note = f"one step: launches from the second-to-last '{marker}' launch up to the last one"
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
    print(f"# {path}: {len(rows)} launches ({note}), {tot / 1e6:.3f} ms total (cold-cache, serialised; compare shares)")
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v / tot * 100:6.2f}%  {c:5d}x  {v / 1e3:10.1f} us  {n}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
