import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
from sonics_b200 import engine as eng
from sonics_b200.sonics import get_alleles
CFG1 = "8|5;9|113;10|89"
CFG2 = "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1"
e = eng.Engine(0)
for name, gts, noise, kw, n in (("cfg1", [CFG1], [0.0], {}, 200000), ("cfg2", [CFG2], [0.0], {}, 200000),
                                ("nocap", [CFG2], [0.0], dict(capture_cycle=40), 200000),
                                ("cfg2big", [CFG2], [0.0], {}, 3000000), ("mixed", [CFG1, CFG2, "3|40;4|25;7|2", "9|1;10|54;11|914;12|15"], [0, 0, 0, 0.0011], {}, 6000)):
    p = eng.make_params(**kw)
    t = e.build_table_from_arrays(p, [get_alleles(g)[0] for g in gts], noise)
    res = {}
    for k in ("persistent", "lockstep"):
        e.set_kernel(k)
        o = e.simulate(p, t, n, seed=5, simpar=True, readout=(n <= 200000))
        e.sync()
        res[k] = {a: b.cpu().numpy() for a, b in o.items()}
    a, b = res["persistent"], res["lockstep"]
    d = a["lnL"] != b["lnL"]
    print(name, "mismatch fraction", d.mean(), "group same", (a["group"] == b["group"]).all(), "simpar same", (a["simpar"] == b["simpar"]).all())
    idx = np.argwhere(d)[:8]
    for g, j in idx:
        print("   g=%d j=%d lnL %.6f vs %.6f  eff=%.5f down=%.5f up=%.5f grp=%d" % (g, j, a["lnL"][g, j], b["lnL"][g, j], a["simpar"][g, j, 3], a["simpar"][g, j, 0], a["simpar"][g, j, 1], a["group"][g, j]))
    if d.any():
        sp = a["simpar"][d]
        print("   mismatching sims: eff mean %.4f (all %.4f) down mean %.4f up mean %.4f" % (sp[:, 3].mean(), a["simpar"][..., 3].mean(), sp[:, 0].mean(), sp[:, 1].mean()))
        rel = np.abs(a["lnL"][d] - b["lnL"][d])
        print("   |dlnL| median %.3g, sentinel flips %d" % (np.median(rel), ((a["lnL"][d] == -999999) != (b["lnL"][d] == -999999)).sum()))
