"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1e.csv profiles/r1_launches_e.txt
    python profiles/summarize.py kernel   gpurun_out/prof_pairs_x3.ncu-rep profiles/r1_ncu_pairs_x3.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    total, n = 0.0, 0
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        agg[name][0] += 1
        agg[name][1] += v
        total += v
        n += 1
    with open(dst, "w") as out:
        out.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        out.write("# source: %s   launches: %d   total: %.3f ms\n" % (src, n, total / 1e6))
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            out.write("%-72s n=%5d total=%9.3f ms avg=%8.1f us share=%5.1f%%\n" % (k[-72:], c, t / 1e6, t / c / 1e3,
                                                                                 100 * t / total))


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write("# ncu --set full --clock-control none; source: %s\n" % src)
        for r in rows[2:]:
            out.write("kernel: %s\n" % r[hdr.index("Kernel Name")])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    out.write("  %-82s %s %s\n" % (k, r[i], units[i]))
    srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(srcp.splitlines()))
    if len(rows) > 2:
        hdr, data = rows[1], rows[2:]
        ia, isrc, iw = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
        hist, tot = collections.Counter(), 0
        for r in data:
            if len(r) <= max(ia, isrc) or not r[ia].isdigit():
                continue
            toks = r[isrc].strip().split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            hist[op] += int(r[ia])
            tot += int(r[ia])
        with open(dst, "a") as out:
            out.write("warp instructions executed: %d; opcode mix (SASS):\n" % tot)
            for op, c in hist.most_common(16):
                out.write("  %-10s %12d %5.1f%%\n" % (op, c, 100.0 * c / tot))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
