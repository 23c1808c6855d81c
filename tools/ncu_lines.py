"""Source-line view of an `ncu --set full --import-source on` capture (kernels compiled with -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--top 40] [--kernel k_units]

Prints, from `ncu --page source --print-source cuda,sass --csv`:
  * the stall mix of the whole kernel (sampled, by stall reason),
  * the source lines with the most stall samples: file:line, samples, share, executed warp instructions,
    lanes active, dominant stall reasons,
  * per source FILE totals.
Each SASS instruction is attributed to the innermost source line the line table gives it (inlined code counts for
the line of the callee).
"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict


def num(v):
    try:
        return int(v.replace(",", ""))
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # several captured launches: keep the file sections whose "Function Name" contains --kernel (default: the first one)
    want = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else None
    keep, func, cur_file, chosen = [], None, None, None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r
            continue
        if len(r) == 2 and r[0] == "Function Name":
            func = r[1]
            if chosen is None and (want is None or (want + "<") in func or (want + "(") in func):
                chosen = func
            if func == chosen:
                keep.append(cur_file)
            continue
        if func == chosen and chosen is not None:
            keep.append(r)
    print("kernel: %s" % chosen)
    rows = keep
    fname, hdr = None, None
    lines = {}          # (file, line) -> dict
    seen_addr = set()
    stall_names = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = os.path.basename(r[1])
            continue
        if len(r) == 2:
            continue
        if r and r[0] == "Line No":
            hdr = r
            stall_names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        # header has "Source" twice (source text, SASS text): csv dict keeps the last; use positions
        if r[0] != "":
            cur = (fname, int(r[0]), r[1].strip())
            continue
        addr = r[2]
        if addr in seen_addr:      # an instruction is listed under every file whose lines it maps to; keep the first
            continue
        seen_addr.add(addr)
        a = lines.setdefault(cur[:2], {"src": cur[2], "samples": 0, "inst": 0, "tinst": 0, "st": defaultdict(int), "n": 0,
                                       "ldl": 0})
        a["samples"] += num(d["# Samples"])
        a["inst"] += num(d["Instructions Executed"])
        a["tinst"] += num(d["Thread Instructions Executed"])
        a["n"] += 1
        if "LDL" in r[3] or "STL" in r[3]:
            a["ldl"] += num(d["Instructions Executed"])
        for s in stall_names:
            a["st"][s] += num(d[s])
    tot_s = sum(a["samples"] for a in lines.values()) or 1
    tot_i = sum(a["inst"] for a in lines.values()) or 1
    mix = defaultdict(int)
    for a in lines.values():
        for s, v in a["st"].items():
            mix[s] += v
    print("total samples %d, executed warp instructions %.4e, SASS instructions %d, local ld/st executed %.3e (%.1f %%)" %
          (tot_s, tot_i, len(seen_addr), sum(a["ldl"] for a in lines.values()),
           100.0 * sum(a["ldl"] for a in lines.values()) / tot_i))
    print("stall mix (share of samples): " + ", ".join("%s %.1f %%" % (s[6:], 100.0 * v / tot_s)
                                                       for s, v in sorted(mix.items(), key=lambda x: -x[1])[:10]))
    files = defaultdict(lambda: [0, 0])
    for (f, l), a in lines.items():
        files[f][0] += a["samples"]; files[f][1] += a["inst"]
    print("\n| file | stall samples | executed warp instructions |\n|---|---|---|")
    for f, v in sorted(files.items(), key=lambda x: -x[1][0]):
        print("| %s | %.1f %% | %.1f %% |" % (f, 100.0 * v[0] / tot_s, 100.0 * v[1] / tot_i))
    print("\n| file:line | samples | instr | lanes | local | top stalls | source |\n|---|---|---|---|---|---|---|")
    for (f, l), a in sorted(lines.items(), key=lambda x: -x[1]["samples"])[:top]:
        st = sorted(a["st"].items(), key=lambda x: -x[1])[:3]
        print("| %s:%d | %.2f %% | %.2f %% | %.1f | %.2f %% | %s | `%s` |" %
              (f, l, 100.0 * a["samples"] / tot_s, 100.0 * a["inst"] / tot_i, a["tinst"] / max(a["inst"], 1),
               100.0 * a["ldl"] / tot_i,
               ", ".join("%s %d%%" % (s[6:], round(100.0 * v / max(a["samples"], 1))) for s, v in st if v),
               a["src"][:90].replace("|", "\\|")))


if __name__ == "__main__":
    main()
