"""Position-dependent clumping (type_of_clumping == 5, evolve_point.F90:484, photonstatistics.f90:176) and Lyman-limit
systems (use_LLS, evolve_point.F90:177-180): oracle self-consistency on the CPU, GPU parity against the oracle."""
import numpy as np
import pytest

import c2ray_b200
from c2ray_b200 import synth
from common import O, frac_err, oracle_grid, oracle_setup, relerr


def _fields(p, seed=11):
    rng = np.random.default_rng(seed)
    n = int(p["mesh"][0])
    clump = np.exp(rng.normal(1.0, 0.8, (n, n, n))).astype(np.float32)          # 1 .. ~30
    tau_cell = 10.0 ** rng.uniform(-3, 0.5, (n, n, n))
    lls = (tau_cell / 6.346e-18).astype(np.float32)                               # column density per cell
    return clump, lls


def test_oracle_hooks_reduce_to_the_scalar_paths():
    p = synth.make_problem(2, n=16, num_src=2, isothermal=False)
    p["NormFlux"] = p["NormFlux"] * 30.0
    oracle_setup(p)
    g = oracle_grid(p)
    base = g.evolve3d(p["dt"]); s0 = g.get_state()
    # a uniform clumping grid equal to the scalar, LLS type 1 with zero column: bitwise the same step
    g = oracle_grid(p)
    g.set_clumping_grid(np.full(16 ** 3, p["clumping"], dtype=np.float32))
    g.set_LLS(1, 0.0)
    same = g.evolve3d(p["dt"]); s1 = g.get_state()
    assert same["niter"] == base["niter"]
    for a, b in zip(s0, s1):
        assert np.array_equal(a, b)
    # real fields change the answer the right way: more recombinations, fewer photons downstream
    clump, lls = _fields(p)
    g = oracle_grid(p); g.set_clumping_grid(clump); g.evolve3d(p["dt"]); s2 = g.get_state()
    assert s2[0][1].sum() < s0[0][1].sum()                  # less ionized hydrogen with clumping > 1
    g = oracle_grid(p); g.set_LLS(2, LLS_grid=lls); g.evolve3d(p["dt"]); s3 = g.get_state()
    assert s3[0][1].sum() < s0[0][1].sum()                  # LLS opacity shields the gas
    # serial and shell order stay bitwise identical with LLS (the LLS term only changes the incoming column)
    res = []
    for order in (0, 1):
        g = oracle_grid(p); g.set_LLS(2, LLS_grid=lls)
        g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"]); g.set_rates_to_zero()
        g.pass_all_sources(order=order)
        res.append(g.get_rates())
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    oracle_grid(p)  # leaves the module state without hooks for the tests that follow


@pytest.mark.gpu
@pytest.mark.parametrize("iso,lls_type", [(False, 2), (True, 1)])
def test_gpu_parity_with_clumping_grid_and_lls(iso, lls_type):
    p = synth.make_problem(2, n=16, num_src=3, isothermal=iso)
    p["NormFlux"] = p["NormFlux"] * 30.0
    tables = oracle_setup(p)
    clump, lls = _fields(p)
    col1 = 0.05 / 6.346e-18
    g = oracle_grid(p)
    g.set_clumping_grid(clump)
    g.set_LLS(lls_type, col1, lls if lls_type == 2 else None)
    c = c2ray_b200.from_problem(p, tables=tables)
    c.set_clumping_grid(clump)
    c.set_LLS(lls_type, col1, lls if lls_type == 2 else None)
    # one source pass: rate grids
    for x in (g, c):
        x.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"]); x.set_rates_to_zero()
    upd_o = g.pass_all_sources(order=1)[0]
    upd_g = c.pass_all_sources(1, p["dt"])
    assert upd_g == upd_o
    for a, b in zip(c.get_rates(), g.get_rates()):
        assert np.array_equal(a != 0, b != 0) and relerr(a, b, 1e-300) < 1e-8
    # the whole step
    g = oracle_grid(p); g.set_clumping_grid(clump); g.set_LLS(lls_type, col1, lls if lls_type == 2 else None)
    so = g.evolve3d(p["dt"])
    c.set_state(p["ndens"], p["xh"], p["xhe"], p["temperature_grid"])
    sg = c.evolve3D(0.0, p["dt"], 0)
    assert sg["niter"] == so["niter"] and list(sg["conv_hist"]) == list(so["conv_hist"])
    xh_o, xhe_o, T_o = g.get_state(); xh, xhe, T = c.get_state()
    assert frac_err(xh, xh_o) < 1 and frac_err(xhe, xhe_o) < 1
    if not iso:
        assert relerr(T, T_o) < 1.3e-7
    tr = g.total_rates(p["dt"], *g.get_work_state()[:2])
    assert relerr([sg["totrec"], sg["totcollisions"], sg["recomions"]], tr, 1e-300) < 1e-8
    # switching the hooks off again restores the plain path
    c.set_clumping_grid(None); c.set_LLS(0)
    g = oracle_grid(p); so = g.evolve3d(p["dt"])
    c.set_state(p["ndens"], p["xh"], p["xhe"], p["temperature_grid"])
    sg = c.evolve3D(0.0, p["dt"], 0)
    assert sg["niter"] == so["niter"] and frac_err(c.get_state()[0], g.get_state()[0]) < 1
    c.close()
    oracle_grid(p)
