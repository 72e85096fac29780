"""Summarise an .ncu-rep here (no GPU): key metrics + SASS opcode histogram + hottest source lines.
usage: python scripts/ncu_read.py gpurun_out/prof_x.ncu-rep [--lines N]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        print("== kernel:", row[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {row[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    h = None
    ops = collections.defaultdict(lambda: [0, 0])
    lines = []
    for r in src:
        if "Source" in r and "# Samples" in r:
            h = r
            continue
        if h is None or len(r) != len(h):
            continue
        try:
            s, e = int(r[h.index("# Samples")]), int(r[h.index("Instructions Executed")])
        except ValueError:
            continue
        text = r[h.index("Source")].strip()
        toks = text.split()
        op = (toks[1] if toks and toks[0].startswith("@") else toks[0] if toks else "?").split(".")[0]
        ops[op][0] += s
        ops[op][1] += e
        lines.append((s, e, text))
    ts = sum(v[0] for v in ops.values()) or 1
    te = sum(v[1] for v in ops.values()) or 1
    print(f"-- SASS opcodes (samples {ts}, warp instructions {te})")
    for op, (s, e) in sorted(ops.items(), key=lambda x: -x[1][1])[:22]:
        print(f"  {op:10s} samples {s / ts * 100:5.1f}%   inst {e / te * 100:5.1f}%  ({e})")
    print("-- hottest SASS lines by stall samples")
    for s, e, text in sorted(lines, key=lambda x: -x[0])[:nlines]:
        print(f"  {s:6d} {e:10d}  {text[:120]}")


if __name__ == "__main__":
    main()
