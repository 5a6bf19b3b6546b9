"""Writes tests/golden/hotpath_small.npz from the CPU oracle (oracle/c2ray_oracle.cpp).

The reference has no golden vectors and cannot be run here, so these fixtures pin the *restatement*:
small seeded inputs and the oracle's outputs for each stage of the path.  tests/test_oracle_cpu.py re-computes them,
tests/test_gpu_golden.py runs the CUDA path on the stored inputs.  Regenerate with: python tools/make_golden.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def inputs():
    rng = np.random.default_rng(20261018)
    n = 96
    lin = 10.0 ** rng.uniform(10, 23, (n, 3)); d = 10.0 ** rng.uniform(9, 21, (n, 3))
    lin[:6] = 0.0
    col6 = np.empty((n, 6)); col6[:, 0::2] = lin; col6[:, 1::2] = lin + d
    return dict(T=10.0 ** np.linspace(0.5, 8.0, 31), col6=col6, vol=10.0 ** rng.uniform(62, 68, n),
                i_state=10.0 ** rng.uniform(-12, 0, n) * 0.99999)


def compute():
    import c2ray_b200
    from oracle import oracle as O
    from common import oracle_setup, oracle_grid
    synth = c2ray_b200.synth
    inp = inputs()
    out = dict(in_T=inp["T"], in_col6=inp["col6"], in_vol=inp["vol"], in_i_state=inp["i_state"])
    p3 = synth.make_problem(3, n=12, num_src=3)
    oracle_setup(p3)
    out["rec_colion"] = np.array([O.rec_colion(t) for t in inp["T"]])
    out["photo_bb_qpl"] = O.photoion_rates_batch(inp["col6"], inp["vol"], [2.0e5, 0.0, 3.0e3], inp["i_state"])
    g = oracle_grid(p3)
    st = g.evolve3d(p3["dt"])
    xh, xhe, T = g.get_state()
    out["frac_xh_cfg3"] = xh; out["frac_xhe_cfg3"] = xhe; out["T_cfg3"] = T.astype(np.float64)
    out["int_cfg3"] = np.array([st["niter"], st["conv_flag"], st["sum_nbox"], st["rt_updates"]], dtype=np.int64)
    out["int_conv_hist_cfg3"] = st["conv_hist"].astype(np.int64)
    out["rates_phih_cfg3"], out["rates_phihe_cfg3"], out["rates_phiheat_cfg3"] = g.get_rates()
    q = synth.make_chemistry_problem(256, seed=11)
    p1 = synth.make_problem(1, n=8)
    oracle_setup(p1)
    ion = np.zeros((256, 15))
    ion[:, 0:2] = q["xh"].T; ion[:, 2:5] = q["xhe"].T; ion[:, 5:7] = q["xh"].T; ion[:, 7:10] = q["xhe"].T
    ion[:, 10:12] = q["xh"].T; ion[:, 12:15] = q["xhe"].T
    phi4 = np.stack([q["phih"], q["phihe"][0], q["phihe"][1], q["phiheat"]], axis=1)
    ri, rT, rn = O.chemistry_batch(q["dt"], q["ndens"], ion, phi4, np.full((256, 3), 1.0e4))
    out["frac_chem"] = ri[:, :10]; out["T_chem"] = rT[:, :2]; out["int_nit_chem"] = rn.astype(np.int64)
    out["photo_bb"] = O.photoion_rates_batch(inp["col6"], inp["vol"], [2.0e5, 0.0, 0.0], inp["i_state"])
    return out


if __name__ == "__main__":
    res = compute()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hotpath_small.npz"), **res)
    print("wrote", {k: v.shape for k, v in res.items()})
