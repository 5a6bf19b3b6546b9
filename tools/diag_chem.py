import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import c2ray_b200
from oracle import oracle as O
from common import oracle_setup
synth = c2ray_b200.synth
for iso in (False, True):
    q = synth.make_chemistry_problem(8192, isothermal=iso)
    p = synth.make_problem(1, n=8, isothermal=iso)
    c = c2ray_b200.from_problem(p, tables=oracle_setup(p))
    n = q["ncells"]
    ion = np.zeros((n, 15))
    ion[:, 0:2] = q["xh"].T; ion[:, 2:5] = q["xhe"].T
    ion[:, 5:7] = q["xh"].T; ion[:, 7:10] = q["xhe"].T
    ion[:, 10:12] = q["xh"].T; ion[:, 12:15] = q["xhe"].T
    phi4 = np.stack([q["phih"], q["phihe"][0], q["phihe"][1], q["phiheat"]], axis=1)
    T3 = np.full((n, 3), 1.0e4)
    gi, gT, gn = c.do_chemistry(q["dt"], q["ndens"], ion, phi4, T3)
    ri, rT, rn = O.chemistry_batch(q["dt"], q["ndens"], ion, phi4, T3)
    names = "h0 h1 he0 he1 he2 hav0 hav1 heav0 heav1 heav2".split()
    print("iso", iso, "nit equal", np.array_equal(gn, rn), "nit max", rn.max(), "T relerr", np.abs(gT/rT-1).max())
    for k in range(10):
        d = np.abs(gi[:, k] - ri[:, k])
        rel = d / np.abs(ri[:, k])
        i = np.argmax(rel)
        big = ri[:, k] > 1e-6
        print(f"  {names[k]:6s} max abs {d.max():.3e}  max rel {rel.max():.3e} at ref={ri[i,k]:.6e} got={gi[i,k]:.6e} | max rel where x>1e-6: {rel[big].max() if big.any() else 0:.3e}  | n(rel>1e-8)={np.sum(rel>1e-8)}")
    c.close()
