"""GPU path against the committed fixtures (tests/golden/hotpath_small.npz, written by tools/make_golden.py from the CPU
oracle; the reference itself ships no vectors), plus size-independent properties at BASELINE's full mesh sizes."""
import os

import numpy as np
import pytest

import c2ray_b200
from oracle import oracle as O
from common import oracle_setup, relerr, frac_err

pytestmark = pytest.mark.gpu
synth = c2ray_b200.synth
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "hotpath_small.npz"))


def _ctx(p):
    return c2ray_b200.from_problem(p, tables=oracle_setup(p))


def test_golden_scalar_kernels():
    p3 = synth.make_problem(3, n=12, num_src=3)
    c = _ctx(p3)
    assert relerr(c.ini_rec_colion_factors(G["in_T"]), G["rec_colion"], 1e-300) < 1e-11
    got = c.photoion_rates(G["in_col6"], G["in_vol"], [2.0e5, 0.0, 3.0e3], G["in_i_state"])
    ref = G["photo_bb_qpl"]
    scale = np.abs(ref).max(axis=0)
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6 * scale)) < 1e-8
    got = c.photoion_rates(G["in_col6"], G["in_vol"], [2.0e5, 0.0, 0.0], G["in_i_state"])
    ref = G["photo_bb"]
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6 * np.abs(ref).max(axis=0))) < 1e-8
    c.close()


def test_golden_chemistry():
    q = synth.make_chemistry_problem(256, seed=11)
    c = _ctx(synth.make_problem(1, n=8))
    ion = np.zeros((256, 15))
    ion[:, 0:2] = q["xh"].T; ion[:, 2:5] = q["xhe"].T; ion[:, 5:7] = q["xh"].T; ion[:, 7:10] = q["xhe"].T
    ion[:, 10:12] = q["xh"].T; ion[:, 12:15] = q["xhe"].T
    phi4 = np.stack([q["phih"], q["phihe"][0], q["phihe"][1], q["phiheat"]], axis=1)
    gi, gT, gn = c.do_chemistry(q["dt"], q["ndens"], ion, phi4, np.full((256, 3), 1.0e4))
    assert np.array_equal(gn, G["int_nit_chem"])
    assert frac_err(gi[:, :10], G["frac_chem"]) < 1
    assert relerr(gT[:, :2], G["T_chem"]) < 1e-8
    c.close()


def test_golden_evolve3d():
    p = synth.make_problem(3, n=12, num_src=3)
    c = _ctx(p)
    st = c.evolve3D(0.0, p["dt"], 0)
    xh, xhe, T = c.get_state()
    assert [st["niter"], st["conv_flag"], st["sum_nbox_all"], st["rt_updates"]] == list(G["int_cfg3"])
    assert list(st["conv_hist"]) == list(G["int_conv_hist_cfg3"][:st["niter"]])
    assert frac_err(xh, G["frac_xh_cfg3"]) < 1 and frac_err(xhe, G["frac_xhe_cfg3"]) < 1
    assert relerr(T, G["T_cfg3"]) < 1.3e-7
    for a, key in zip(c.get_rates(), ("rates_phih_cfg3", "rates_phihe_cfg3", "rates_phiheat_cfg3")):
        b = G[key]
        assert relerr(a, b, 1e-300) < 1e-8, key   # pure relative
    c.close()


# ---- properties at full size (no oracle run needed) -------------------------------------------------------------
def test_full_size_config2_properties():
    """BASELINE configs[1] at 128^3 / 16 sources: one RT pass + one global pass.  Coverage (every cell updated exactly once
    per source), linearity of the rate grids in the source strength, determinism of the integer bookkeeping and the
    1-rank == 2-rank-partition sum."""
    p = synth.make_problem(2, n=128, num_src=16)
    c = c2ray_b200.from_problem(p)
    c.begin_step()
    c.set_rates_to_zero()
    upd = c.pass_all_sources(1, p["dt"])
    assert upd == 16 * 128 ** 3
    r1 = c.get_rates()
    assert all(np.all(np.isfinite(a)) for a in r1) and np.all(r1[0] > 0)
    nb = [c.do_source(p["dt"], ns, 1)[0] for ns in (1, 16)]
    assert nb == [1, 1]  # subboxsize = mesh: one sub-box covers the periodic box
    # linearity: doubling every NormFlux doubles every rate (tables are linear in the flux)
    c.set_sources(p["srcpos"], 2.0 * p["NormFlux"])
    c.set_rates_to_zero()
    c.pass_all_sources(1, p["dt"])
    r2 = c.get_rates()
    for a, b in zip(r2, r1):
        assert relerr(a, 2.0 * b, 1e-9 * np.abs(b).max() + 1e-300) < 1e-9
    # source partition: rank 0 of 2 + rank 1 of 2 == all
    c.set_sources(p["srcpos"], p["NormFlux"])
    parts = []
    for rank in (0, 1):
        c.set_rank(rank, 2)
        c.set_rates_to_zero()
        u = c.pass_all_sources(1, p["dt"])
        assert u == 8 * 128 ** 3
        parts.append(c.get_rates())
    for k in range(3):
        assert relerr(parts[0][k] + parts[1][k], r1[k], 1e-9 * np.abs(r1[k]).max() + 1e-300) < 1e-10
    c.set_rank(0, 1)
    # global pass invariants
    c.set_rates(*r1)
    cf = c.global_pass(p["dt"])
    xh_av, xhe_av, xh_i, xhe_i = c.get_work_state()
    assert 0 < cf <= 128 ** 3
    assert np.abs(xh_i.sum(axis=0) - 1).max() < 1e-12 and np.abs(xhe_i.sum(axis=0) - 1).max() < 1e-12
    assert xh_i.min() >= 1e-20 and xhe_i.min() >= 0.999e-20  # He floors are renormalised (doric.f90:254-257)
    T = c.get_state()[2]
    assert np.all(T[0] >= 1.0) and np.all(np.isfinite(T))
    c.close()


def test_full_size_chemistry_idempotent_when_dark():
    """256^3 cells with zero rates and neutral gas: the chemistry must leave the (epsilon-floored) state neutral, vote
    'converged' everywhere and take exactly one do_chemistry iteration per cell (isothermal)."""
    n = 256
    p = synth.make_problem(1, n=n, isothermal=True)
    c = c2ray_b200.from_problem(p)
    c.begin_step()
    c.set_rates(np.zeros((n, n, n)), np.zeros((2, n, n, n)))
    cf, nit = c.global_pass(p["dt"], want_nit=True)
    assert cf == 0 and nit.min() == 1 and nit.max() <= 2
    xh_av, xhe_av, xh_i, xhe_i = c.get_work_state()
    assert xh_i[1].max() < 1e-4 and xhe_i[2].max() < 1e-10
    c.close()


def test_full_size_config3_properties():
    """BASELINE configs[2] at full size (256^3, 1000 sources, BB + QPL on the 50 brightest, subboxsize 10): properties that
    need no oracle run -- repeatability of the integer bookkeeping, linearity of the rate grids in the source strengths
    (the sub-box termination test is scale invariant), and the 2-rank source partition summing to the single-rank grids."""
    p = synth.make_problem(3)
    assert p["mesh"][0] == 256 and len(p["NormFlux"]) == 1000 and int((p["NormFluxQPL"] > 0).sum()) == 50
    c = c2ray_b200.from_problem(p)
    c.begin_step()
    c.set_rates_to_zero()
    upd = c.pass_all_sources(1, p["dt"])
    r1 = c.get_rates()
    assert upd > 1000 * 21 ** 3 and upd < 1000 * 256 ** 3    # every source traces at least its first sub-box
    assert all(np.all(np.isfinite(a)) and a.min() >= 0 for a in r1) and r1[0].max() > 0 and r1[2].max() > 0
    c.set_rates_to_zero()
    assert c.pass_all_sources(1, p["dt"]) == upd
    for a, b in zip(c.get_rates(), r1):
        assert relerr(a, b, 1e-9 * np.abs(b).max() + 1e-300) < 1e-9      # atomics: summation order only
    c.set_sources(p["srcpos"], 2.0 * p["NormFlux"], None, 2.0 * p["NormFluxQPL"])
    c.set_rates_to_zero()
    assert c.pass_all_sources(1, p["dt"]) == upd
    for a, b in zip(c.get_rates(), r1):
        assert relerr(a, 2.0 * b, 1e-9 * np.abs(b).max() + 1e-300) < 1e-9
    c.set_sources(p["srcpos"], p["NormFlux"], None, p["NormFluxQPL"])
    parts, upds = [], 0
    for rank in (0, 1):
        c.set_rank(rank, 2)
        c.set_rates_to_zero()
        upds += c.pass_all_sources(1, p["dt"])
        parts.append(c.get_rates())
    assert upds == upd
    for k in range(3):
        assert relerr(parts[0][k] + parts[1][k], r1[k], 1e-9 * np.abs(r1[k]).max() + 1e-300) < 1e-9
    c.close()
