"""Seeded reservoir parameter sets that visit every branch of every rule (shared by tests/test_gpu_route.py and
tools/bench_route.py)."""
import numpy as np
from vicres_b200 import route_abi


def random_reservoir(rng, rule, ndays):
    """parameters spread so that every branch of the rule is visited over a few hundred days"""
    vlive = float(rng.uniform(3000, 30000)); vdead = float(vlive * rng.uniform(0.1, 0.3))
    hmax = float(rng.uniform(110, 160)); hmin = float(hmax - rng.uniform(15, 40))
    q = float(rng.uniform(5, 60))
    d = route_abi.reservoir(hresermax=hmax, hresermin=hmin, v_live=vlive, v_dead=vdead, head=float(0.6 * (hmax - hmin) + 20),
                            q_design=q, ope_year=int(rng.choice([1999, 1999, 2001])), v_initial=float(vdead + rng.uniform(0.1, 0.9) * vlive),
                            seepage=float(rng.choice([0.0, 0.05])), infiltration=float(rng.choice([0.0, 0.02])),
                            hydropower_type=int(rng.choice([0, 1, 1, 1])), hydropower_generator=int(rng.choice([0, 1, 1])),
                            env_flow=float(rng.choice([0.05, 0.1])), rule=rule, bathymetry=0)
    if rule == 1:
        t = sorted(int(x) for x in rng.choice(np.arange(20, 340), 2, replace=False))
        d.update(hmax=hmax - 1, hmin=hmin + 2, op1=[t[1], t[0]] if rng.uniform() < 0.5 else [t[0], t[1]])
    elif rule == 2:
        d.update(reslv=[float(x) for x in rng.uniform(hmin, hmax, 12)])
    elif rule == 3:
        d.update(demand=float(q * 0.3), x1=float(rng.uniform(0.1, 1.2)), x2=float(vdead + 0.3 * vlive), x3=float(vdead + 0.8 * vlive),
                 x4=float(rng.uniform(0.1, 1.4)))
    elif rule == 5:
        d.update(demand5=[float(q * rng.uniform(0.1, 0.5)) for _ in range(12)], op5x1=[float(rng.uniform(0.1, 1.2)) for _ in range(12)],
                 op5x2=[float(vdead + rng.uniform(0.2, 0.4) * vlive) for _ in range(12)],
                 op5x3=[float(vdead + rng.uniform(0.7, 0.9) * vlive) for _ in range(12)],
                 op5x4=[float(rng.uniform(0.1, 1.4)) for _ in range(12)])
    elif rule == 4:
        d.update(series=rng.uniform(0, 1.5 * q, ndays).astype(np.float32))
    elif rule == 6:
        d.update(series=rng.uniform(vdead, vdead + vlive, ndays).astype(np.float32))
    if rng.uniform() < 0.3:
        d.update(irrigation=rng.uniform(0, 0.2 * q, ndays).astype(np.float32))
    return d
