"""Turn ncu exports into the small, committed summaries under profiles/.

    python profiles/summarize_ncu.py full   <raw.csv>      <out.md>   # `ncu -i X.ncu-rep --page raw --csv`
    python profiles/summarize_ncu.py launch <launches.csv> <out.md>   # `--metrics gpu__time_duration.sum` list

`full`: one row per profiled launch with the metrics the roofline discussion in DESIGN.md uses.
`launch`: per-kernel launch count, total and mean device time, share of the summed kernel time.
"""
import csv
import sys
from collections import OrderedDict

FULL = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64 %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall LG"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall noinst"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall bar"),
]


def short(name):
    name = name.replace("<unnamed>::", "").replace("void ", "")
    return name.split("(")[0]


def fmt(v, unit):
    try:
        f = float(v)
    except ValueError:
        return v
    if unit in ("byte", "Kbyte", "Mbyte", "Gbyte"):
        return "%.2f %s" % (f, unit.replace("byte", "B"))
    if unit in ("ns", "us", "ms", "usecond", "nsecond", "msecond"):
        return "%.2f %s" % (f, unit[:2])
    if f == int(f) and abs(f) < 1e12:
        return "%d" % f
    return "%.2f" % f


def full(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [(m, t) for m, t in FULL if m in ix]
    with open(dst, "w") as out:
        out.write("| kernel | grid | block | " + " | ".join(t for _, t in cols) + " |\n")
        out.write("|---|---|---|" + "---|" * len(cols) + "\n")
        for r in rows[2:]:
            out.write("| %s | %s | %s | " % (short(r[ix["Kernel Name"]]), r[ix["Grid Size"]], r[ix["Block Size"]]))
            out.write(" | ".join(fmt(r[ix[m]], units[ix[m]]) for m, _ in cols) + " |\n")
        out.write("\nColumns: ncu `--set full --clock-control none`, one row per profiled launch (caches flushed "
                  "between replays, so dram wr under-reports stores that are still dirty in the 126 MB L2).\n")


def launch(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        try:
            ns = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as out:
        out.write("| kernel | launches | total us | mean us | share of kernel time |\n|---|---|---|---|---|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write("| %s | %d | %.1f | %.2f | %.1f %% |\n" % (k, n, ns / 1e3, ns / n / 1e3, 100.0 * ns / total))
        out.write("\nncu `--metrics gpu__time_duration.sum --clock-control none`: cold-cache, serialised launches; "
                  "compare shares, not absolutes.\n")


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2], sys.argv[3])
