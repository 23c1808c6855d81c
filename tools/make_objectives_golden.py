"""Golden vectors for the calibration objectives (NSE, TRMSE, MSDE, ROCE), made by executing THE REFERENCE'S OWN
functions: toolbox/calibration/basin_calibration_eNSGAII.py cannot be imported (it needs Platypus and runs its
optimisation at import time), so the four `calculate_*` function definitions are cut out of its source with
`ast` and executed here, unmodified.  Output: tests/golden/objectives.npz."""
import ast
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/toolbox/calibration/basin_calibration_eNSGAII.py"


def reference_functions():
    tree = ast.parse(open(SRC).read())
    want = {"calculate_nse", "calculate_trmse", "calculate_msde", "calculate_roce"}
    mod = ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want], type_ignores=[])
    ns = {"np": np}
    exec(compile(mod, SRC, "exec"), ns)
    return ns


def main():
    f = reference_functions()
    rng = np.random.default_rng(77)
    ndays, nsets = 2130, 24                      # eFAST_setup.txt horizon
    obs = np.abs(rng.gamma(0.7, 300.0, ndays) * (1 + 0.5 * np.sin(np.arange(ndays) * 2 * np.pi / 365.0)))
    sim = (obs[None, :] * rng.lognormal(0.0, 0.4, (nsets, ndays)) + rng.normal(0, 20.0, (nsets, ndays))).astype(np.float32)
    sim[3] = 0.0
    sim[5, ::7] *= -1.0                           # negative simulated flows (the toolbox takes abs)
    res = np.zeros((2, nsets, 4))
    for h, handle in enumerate([True, False]):
        for s in range(nsets):
            x = sim[s].astype(np.float64)
            res[h, s] = [f["calculate_nse"](obs, x, handle), f["calculate_trmse"](obs, x, handle),
                         f["calculate_msde"](obs, x, handle), f["calculate_roce"](obs, x, handle)]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "objectives.npz"), obs=obs, sim=sim, ref=res)
    print(res[0, :3])


if __name__ == "__main__":
    main()
