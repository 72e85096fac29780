"""Stall samples of an .ncu-rep by category and by 40-instruction window of the SASS stream (no GPU needed).
usage: python scripts/ncu_stalls.py gpurun_out/prof_x.ncu-rep [window]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    win = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(r for r in rows if "Source" in r and "# Samples" in r)
    idx = {n: i for i, n in enumerate(h)}
    cats = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tot = collections.Counter()
    data = []
    for r in rows:
        if len(r) != len(h):
            continue
        try:
            s = int(r[idx["# Samples"]])
        except ValueError:
            continue
        d = {c: int(r[idx[c]] or 0) for c in cats}
        for c in cats:
            tot[c] += d[c]
        data.append((r[idx["Source"]].strip(), s, int(r[idx["Instructions Executed"]]), d))
    total = sum(tot.values())
    print("total samples", total)
    for c, v in tot.most_common():
        print(f"{c:28s} {v:7d} {100 * v / total:5.1f}%")
    print(f"\n-- SASS stream in address order, samples per {win}-instruction window")
    for i in range(0, len(data), win):
        w = data[i:i + win]
        s = sum(x[1] for x in w)
        if s < total / 300:
            continue
        cc = collections.Counter()
        for x in w:
            for c, v in x[3].items():
                cc[c] += v
        ex = sum(x[2] for x in w) / len(w)
        top = ", ".join(f"{c[6:]}:{v}" for c, v in cc.most_common(4))
        ops = collections.Counter((x[0].split()[1] if x[0].startswith("@") else x[0].split()[0]).split(".")[0] for x in w if x[0])
        print(f"{i:5d} {s:6d} ({100 * s / total:4.1f}%) exec/instr {ex:9.0f}  {top} | " + " ".join(f"{o}{n}" for o, n in ops.most_common(5)))


if __name__ == "__main__":
    main()
