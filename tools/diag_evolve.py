"""Where does the GPU differ from the oracle after a full evolve3D?  (three-SED 20^3 case of tests/test_gpu_extra.py)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, c2ray_b200
from oracle import oracle as O
from common import oracle_setup, oracle_grid
synth = c2ray_b200.synth
p = synth.make_problem(3, n=20, num_src=4)
p["pl"] = dict(index=2.5, minfreq=p["qpl"]["minfreq"] * 0.2, maxfreq=p["qpl"]["maxfreq"], S_star=1e48)
p["NormFluxPL"] = np.array([0.0, 2.0e3, 0.0, 5.0e2]); p["NormFluxQPL"] = np.array([1.0e3, 0.0, 0.0, 3.0e2]); p["NormFlux"] = np.array([3.0e6, 0.0, 2.0e6, 1.0e6])
tables = oracle_setup(p)
g = oracle_grid(p)
c = c2ray_b200.from_problem(p, tables=tables)
# iteration by iteration, each side on its own state
g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"]); c.begin_step()
for it in range(1, 8):
    g.set_rates_to_zero(); g.pass_all_sources(); c.set_rates_to_zero(); c.pass_all_sources(it, p["dt"])
    ro, rg = g.get_rates(), c.get_rates()
    rr = max(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6 * np.abs(b).max())) for a, b in zip(rg, ro))
    cfo = g.global_pass(p["dt"]); cfg = c.global_pass(p["dt"])
    wo, wg = g.get_work_state(), c.get_work_state()
    names = ("xh_av", "xhe_av", "xh_int", "xhe_int")
    msg = []
    for nme, a, b in zip(names, wg, wo):
        d = np.abs(a - b); m = d / (1e-8 * np.abs(b) + 2e-10); i = np.unravel_index(np.argmax(m), m.shape)
        msg.append(f"{nme}: mixed {m.max():.2f} at x={b[i]:.3e} abs {d[i]:.2e} comp {i[0]}")
    To, Tg = g.get_state()[2], c.get_state()[2]
    print(f"it {it}: rates rel {rr:.1e} conv {cfo}/{cfg} T rel {np.abs(Tg[:2].astype(float)/To[:2]-1).max():.1e} | " + " | ".join(msg))
c.close()
