"""Run one oracle case in a fresh process (the oracle library keeps module-global state; C2RAY_ORACLE_LIB picks the
build variant) and save the results as .npz.  Used by tests/test_oracle_cpu.py for the arithmetic noise-floor check."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import c2ray_b200
from oracle import oracle as O
from common import oracle_setup, oracle_grid
synth = c2ray_b200.synth

def main(case, out):
    if case == "chem":
        res = {}
        for iso in (False, True):
            q = synth.make_chemistry_problem(8192, isothermal=iso)
            p = synth.make_problem(1, n=8, isothermal=iso)
            oracle_setup(p)
            n = q["ncells"]
            ion = np.zeros((n, 15))
            ion[:, 0:2] = q["xh"].T; ion[:, 2:5] = q["xhe"].T; ion[:, 5:7] = q["xh"].T; ion[:, 7:10] = q["xhe"].T
            ion[:, 10:12] = q["xh"].T; ion[:, 12:15] = q["xhe"].T
            phi4 = np.stack([q["phih"], q["phihe"][0], q["phihe"][1], q["phiheat"]], axis=1)
            ri, rT, rn = O.chemistry_batch(q["dt"], q["ndens"], ion, phi4, np.full((n, 3), 1.0e4))
            res[f"ion_{int(iso)}"] = ri; res[f"T_{int(iso)}"] = rT; res[f"nit_{int(iso)}"] = rn
        np.savez(out, **res)
    elif case == "evolve":
        p = synth.make_problem(1, n=20)
        oracle_setup(p)
        g = oracle_grid(p)
        st = g.evolve3d(p["dt"])
        xh, xhe, T = g.get_state()
        np.savez(out, xh=xh, xhe=xhe, T=T, niter=st["niter"], conv_hist=st["conv_hist"])

main(sys.argv[1], sys.argv[2])
