"""Summarise an `ncu --set full --import-source on` capture of the step kernel (k_units) into markdown + JSON.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "<label>" <cells> <records> [--traffic-json profiles/k_step_traffic.json]

Prints the key counters, the stall mix and a per-device-function table (share of executed warp
instructions, lane utilisation, share of stall samples), mapping SASS addresses to the out-of-line
device functions through the symbol table of vicres_b200/libvicres.so.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "smsp__cycles_elapsed.avg",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, label, cells, recs = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    print("## %s\n" % label)
    print("capture: `%s`, %d cells x %d records in the profiled launch\n" % (os.path.basename(rep), cells, recs))
    print("| counter | value |\n|---|---|")
    for k in KEYS:
        if k in m:
            print("| `%s` | %s %s |" % (k, m[k][0], m[k][1]))
    stalls = sorted(((float(v[0].replace(",", "")), h) for h, v in m.items()
                     if "average_warps_issue_stalled" in h and "not_issued" not in h and v[0]), reverse=True)
    print("\nstall mix (warps per issue-active cycle): " +
          ", ".join("%s %.2f" % (h.split("issue_stalled_")[1].split("_per_")[0], x) for x, h in stalls[:8]))
    dram = to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])
    per = dram / (cells * recs)
    print("\nDRAM traffic: %.3f GB = **%.0f B per cell-timestep** (algorithmic 280 B)\n" % (dram / 1e9, per))
    if "--traffic-json" in sys.argv:
        path = sys.argv[sys.argv.index("--traffic-json") + 1]
        def cnt(k):
            return float(m[k][0].replace(",", "")) if k in m else 0.0
        # thread-level FP64 arithmetic instructions of the launch: (sum over SMSPs per elapsed cycle) x cycles
        fp64 = sum(cnt("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % o)
                   for o in ("dfma", "dmul", "dadd")) * cnt("smsp__cycles_elapsed.avg")
        json.dump({"dram_bytes_per_cell_step": per, "fp64_thread_inst_per_cell_step": fp64 / (cells * recs),
                   "source": os.path.basename(rep), "cells": cells, "records": recs,
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum, and DFMA+DMUL+DADD thread instructions, of one k_units launch / (cells x records)"},
                  open(path, "w"), indent=1)
    if "--no-functions" in sys.argv:     # the symbol table of the current libvicres.so does not match an old capture
        return
    # per-function table
    so = os.path.join(ROOT, "vicres_b200", "libvicres.so")
    elf = subprocess.run(["cuobjdump", "-elf", so], capture_output=True, text=True).stdout
    syms = []
    for l in elf.splitlines():
        p = l.split()
        # the default kernel instance only (k_units<0>): its out-of-line device functions are "$<kernel>$<function>"
        if len(p) < 7 or "k_unitsILi0E" not in l:
            continue
        try:
            val, sz = int(p[1], 16), int(p[2], 16)
        except ValueError:
            continue
        if sz == 0:
            continue
        name = re.sub(r"\$?_ZN\d+_GLOBAL__N__[0-9a-f_]*vicres_capi_cu_[0-9a-f]*7k_unitsILi0EEEvNS_8DevModelENS_10DevForcingEPd", "", p[-1])
        syms.append((val, sz, name))
    subs = sorted(s for s in syms if s[2].startswith("$"))
    src = ncu_csv(rep, "source")
    h = src[1]
    ia, ii, it, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    base = int(src[2][ia], 16)
    agg = {}
    for r in src[2:]:
        off = int(r[ia], 16) - base
        name = "k_units (kernel body: loads, stores, record loop)"
        for v, sz, n in subs:
            if v <= off < v + sz:
                name = n
                break
        a = agg.setdefault(name, [0, 0, 0, 0])
        a[0] += int(r[ii]); a[1] += int(r[it]); a[2] += int(r[isamp]); a[3] += 1
    tot, tots = sum(a[0] for a in agg.values()), sum(a[2] for a in agg.values())
    def demangle(n):
        m = re.match(r"^\$_ZN2vr(\d+)", n)
        if not m:
            return n
        k = int(m.group(1))
        return "vr::" + n[m.end():m.end() + k]
    print("| device function | SASS instrs | share of executed warp-instrs | lanes active | share of stall samples |\n|---|---|---|---|---|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][2])[:16]:
        print("| %s | %d | %.1f %% | %.1f / 32 | %.1f %% |" % (demangle(n), a[3], 100 * a[0] / tot, a[1] / max(a[0], 1), 100 * a[2] / tots))
    print("\ntotal executed warp instructions %.3e; thread instructions per cell-timestep %.0f\n" %
          (tot, sum(a[1] for a in agg.values()) / max(1, cells * recs)))


if __name__ == "__main__":
    main()
