"""Summarise an `ncu --metrics ... --csv --log-file x.csv` launch list: per kernel the number of launches, total and mean
duration, share of the listed GPU time, and the mean of every other metric that was collected.

    python tools/ncu_launches.py gpurun_out/launches.csv
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = defaultdict(lambda: defaultdict(list))
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
        try:
            per[name][r[im]].append(float(r[iv].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(m["gpu__time_duration.sum"]) for m in per.values())
    others = sorted({k for m in per.values() for k in m} - {"gpu__time_duration.sum"})
    short = {"launch__registers_per_thread": "regs", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps %",
             "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue %",
             "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64 %",
             "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes", "smsp__inst_executed.sum": "warp inst"}
    print("| kernel | launches | total ms | share | mean us | " + " | ".join(short.get(o, o) for o in others) + " |")
    print("|---|---|---|---|---|" + "---|" * len(others))
    for name, m in sorted(per.items(), key=lambda x: -sum(x[1]["gpu__time_duration.sum"])):
        d = m["gpu__time_duration.sum"]
        cells = []
        for o in others:
            v = m.get(o, [])
            cells.append("%.4g" % (sum(v) / len(v)) if v else "")
        print("| `%s` | %d | %.2f | %.1f %% | %.1f | %s |" % (name, len(d), sum(d) / 1e6, 100 * sum(d) / tot, sum(d) / len(d) / 1e3,
                                                           " | ".join(cells)))


if __name__ == "__main__":
    main()
