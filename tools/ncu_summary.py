"""Summarise an ncu report (`ncu --set full`) and/or a launch list (`--metrics gpu__time_duration.sum --csv`)
into small text files that are committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_x_ncu.md
    python tools/ncu_summary.py --launches gpurun_out/launches.csv profiles/r1_x_launches.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % (active)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
]


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0][:60]


def report(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full summary of `%s`" % rep, "",
             "(cold-cache, serialised replays: compare shares and ratios, not absolute times)", ""]
    lines.append("| metric | unit | " + " | ".join(short(r[idx["Kernel Name"]]) for r in data) + " |")
    lines.append("|---|---|" + "---|" * len(data))
    for key, label in METRICS:
        if key in idx:
            lines.append("| %s | %s | " % (label, units[idx[key]]) + " | ".join(r[idx[key]] for r in data) + " |")
    open(out, "w").write("\n".join(lines) + "\n")


def launches(path, out):
    txt = open(path).read()
    start = txt.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v_us = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v_us
    tot = sum(a[1] for a in agg.values())
    lines = ["# kernel launch list `%s` (ncu --metrics gpu__time_duration.sum --clock-control none)" % path, "",
             "| kernel | launches | total us | mean us | share |", "|---|---|---|---|---|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| %s | %d | %.1f | %.2f | %.1f%% |" % (k, c, t, t / c, 100 * t / tot))
    lines.append("")
    lines.append("total launches %d, total device time %.1f us" % (sum(a[0] for a in agg.values()), tot))
    open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        report(sys.argv[1], sys.argv[2])
