"""Where do the GPU and the oracle part ways?  Runs a seeded grid on the GPU and through the oracle and reports, for the
cells that end up outside the parity envelope, the first record and variables that differ by more than 1e-9 and what
happens next (run on the GPU box; diagnostic for profiles/parity_r2.md).

    python tools/parity_probe.py [cfg2|cfg3] [cells] [records]
"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import oracle_lib
from vicres_b200 import abi, model, synth_soa
from common import relerr, RTOL, ENV_ABS, ENV_REL
from test_gpu_parity import _oracle_threads

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
if cfg == "cfg2":
    C, nrec, nb, dt, cold, kw = 2000, 3653, 1, 24, 0.0, {}
else:
    C, nrec, nb, dt, cold, kw = 500, 2920, 5, 3, 5.0, {"lat_range": (35.0, 65.0), "overstory_frac": 0.3}
if len(sys.argv) > 2:
    C = int(sys.argv[2])
if len(sys.argv) > 3:
    nrec = int(sys.argv[3])
lib, cells = synth_soa.make_cells(C, nlayer=3, nbands=nb, seed=20261019, **kw)
f, month, doy = synth_soa.make_forcing(cells, nrec, dt_hours=dt, seed=7, cold=cold)
names = list(abi.OUT_NAMES)
opt = abi.default_options(3, nb, dt, out=names)
m = model.VicRes(opt, lib, cells)
m.init_state(f[0, 0])
out, status = m.step(month, doy, f)
ref = _oracle_threads(opt, lib, cells, month, doy, f)
e = relerr(out, ref)
a = np.abs(out - ref)
inside = (e <= RTOL) | (a <= ENV_ABS + ENV_REL * np.maximum(np.abs(out), np.abs(ref)))
bad = np.where((~inside).any(axis=(0, 1)))[0]
print("%s: %d cells x %d records, %.6f of values within 1e-9, %d cells (%.4f) leave the envelope" %
      (cfg, C, nrec, (e <= RTOL).mean(), len(bad), len(bad) / C))
first_vars = {}
for c in bad[:400]:
    ec = e[:, :, c]
    r_first = int(np.argmax((ec > RTOL).any(axis=0)))
    v_first = [names[v] for v in np.where(ec[:, r_first] > RTOL)[0]]
    r_env = int(np.argmax((~inside[:, :, c]).any(axis=0)))
    v_env = [names[v] for v in np.where(~inside[:, r_env, c])[0]]
    key = ",".join(v_first[:4])
    first_vars[key] = first_vars.get(key, 0) + 1
    if len(first_vars) <= 12 and first_vars[key] <= 2:
        i = names.index(v_first[0])
        print("cell %d: first >1e-9 at rec %d in %s (e.g. %s %.17g vs %.17g, rel %.2e); leaves the envelope at rec %d in %s; "
              "SWE there %.6g vs %.6g, WDEW %.3g vs %.3g, air temp %.3f" %
              (c, r_first, v_first, v_first[0], out[i, r_first, c], ref[i, r_first, c], ec[i, r_first], r_env, v_env[:5],
               out[names.index("SWE"), r_env, c], ref[names.index("SWE"), r_env, c], out[names.index("WDEW"), r_env, c],
               ref[names.index("WDEW"), r_env, c], f[abi.VR_F_AIR_TEMP if hasattr(abi, "VR_F_AIR_TEMP") else 0, r_env, c]))
print("first differing variables (cells):", sorted(first_vars.items(), key=lambda x: -x[1])[:12])
# the population of small mismatches (inside the envelope): by variable
cnt = (e > RTOL).sum(axis=(1, 2))
print("values beyond 1e-9 by variable:", {names[v]: int(cnt[v]) for v in np.argsort(-cnt)[:12] if cnt[v]})
