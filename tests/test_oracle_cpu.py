"""CPU tests of the oracle itself (no GPU): literal semantics, an independent numpy restatement of the scalar kernels,
analytic limits and invariants, sweep-order equivalence, the reference's sub-box bookkeeping, golden fixtures.
The reference ships no tests or golden vectors: these checks are what pins the restatement."""
import os
import subprocess
import sys

import numpy as np
import pytest

import c2ray_b200
from oracle import oracle as O
from common import oracle_setup, oracle_grid, relerr, frac_err

synth = c2ray_b200.synth
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = lambda x: float(np.float32(x))


# ---------------------------------------------------------------------------------------------------------------
# independent numpy restatement of cgsconstants.f90:140-266 and doric.f90:35-313 (written from the Fortran, not from
# oracle/c2ray_oracle.cpp), with the default-real literal rule applied by hand
# ---------------------------------------------------------------------------------------------------------------
EV2K = float(np.float32(1.0) / np.float32(8.617e-05))
ETH0, ETHE0, ETHE1 = f32(13.598), f32(24.587), f32(54.416)
TEMPH0, TEMPHE0, TEMPHE1 = ETH0 * EV2K, ETHE0 * EV2K, ETHE1 * EV2K
COLH0 = f32(1.3e-8) * f32(0.83) * 1.0 / (ETH0 * ETH0)
COLHE0 = f32(1.3e-8) * f32(0.63) * 2.0 / (ETHE0 * ETHE0)
COLHE1 = f32(1.3e-8) * f32(1.30) * 1.0 / (ETHE1 * ETHE1)
ABU_HE = f32(0.074)
EPS = 1e-20


def np_rec_colion(T):
    lam = 2.0 * (TEMPH0 / T)
    arech0 = f32(1.269e-13) * lam ** 1.503 / (1.0 + (lam / f32(0.522)) ** f32(0.470)) ** f32(1.923)
    brech0 = f32(2.753e-14) * lam ** 1.500 / (1.0 + (lam / f32(2.740)) ** f32(0.407)) ** f32(2.242)
    if T < 9.0e3:
        areche0 = 1.269e-13 * lam ** 1.503 / (1.0 + (lam / f32(0.522)) ** f32(0.470)) ** f32(1.923)
        breche0 = 2.753e-14 * lam ** 1.500 / (1.0 + (lam / f32(2.740)) ** f32(0.407)) ** f32(2.242)
    else:
        l0 = 2.0 * (TEMPHE0 / T)
        diel = 1.9e-3 * T ** (-1.5) * np.exp(-4.7e5 / T) * (1.0 + 0.3 * np.exp(-9.4e4 / T))
        areche0 = 3.000e-14 * l0 ** 0.654 + diel
        breche0 = 1.260e-14 * l0 ** 0.750 + diel
    l1 = 2.0 * (TEMPHE1 / T)
    breche1 = 5.5060e-14 * l1 ** 1.5 / (1.0 + (l1 / 2.740) ** 0.407) ** 2.242
    areche1 = f32(2.538e-13) * l1 ** 1.503 / (1.0 + (l1 / 0.522) ** 0.470) ** 1.923
    treche1 = 3.4e-13 * (T / 1.0e4) ** (-0.6)
    v = 0.285 * (T / 1.0e4) ** 0.119
    sq = np.sqrt(T)
    return np.array([arech0, brech0, areche0, breche0, areche0 - breche0, areche1, breche1, treche1,
                     COLH0 * sq * np.exp(-TEMPH0 / T), COLHE0 * sq * np.exp(-TEMPHE0 / T), COLHE1 * sq * np.exp(-TEMPHE1 / T), v])


def np_doric(dt, rhe, ion, phi, fr, T, clumping=1.0):
    """ion: dict with h, he, h_old, he_old ; returns (h, he, h_av, he_av)"""
    (arech0, brech0, areche0, breche0, oreche0, areche1, breche1, treche1, cHI, cHeI, cHeII, v) = np_rec_colion(T)
    yfrac, zfrac, y2a, y2b = fr
    pfrac = 0.96
    hef = ABU_HE / (1.0 - ABU_HE)
    ffrac = max(min(10.0 * ion["h"][0], 1.0), 0.01)
    wfrac = (1.425 - 0.737) + 0.737 * yfrac
    ahB, ahe1, aheB, aheA = clumping * brech0, clumping * oreche0, clumping * breche0, clumping * areche0
    ahe2B, ahe2A, ahe22 = clumping * breche1, clumping * areche1, clumping * treche1
    ahe21 = ahe2A - ahe2B
    aih0 = max(phi[0] + rhe * cHI, 1e-200)
    aihe0 = max(phi[1] + rhe * cHeI, 1e-200)
    aihe1 = max(phi[2] + rhe * cHeII, 1e-200)
    L = -(aih0 + rhe * ahB)
    M = (yfrac * rhe * ahe1 + pfrac * rhe * aheB) * hef
    N = ((ffrac * zfrac * (1.0 - v) + v * wfrac) * ahe2B + ahe22 + (1.0 - y2a - y2b) * ahe21) * hef * rhe
    P = -aihe0 - aihe1 - rhe * (aheA - (1.0 - yfrac) * ahe1)
    E = -rhe * (ahe2A - y2a * ahe21)
    Q = -aihe0 + rhe * ahe2B * (ffrac * (1.0 - zfrac) * (1.0 - v) + v * (1.425 - wfrac)) - E + ahe21 * y2b * rhe
    A = np.array([[L, M, N], [0.0, P, Q], [0.0, aihe1, E]])
    g = np.array([aih0, aihe0, 0.0])
    x0 = np.array([ion["h_old"][1], ion["he_old"][1], ion["he_old"][2]])
    return A, g, x0


def expm_solution(A, g, x0, dt):
    """x(dt) and its time average for dx/dt = A x + g, by a 60-digit matrix exponential (mpmath): x = r + e^{A dt}(x0-r),
    <x> = r + A^-1 (e^{A dt} - I)(x0-r)/dt with r = -A^-1 g.  The system is stiff (rates from 1e-27 to 1e-9 /s), so a
    double-precision eigen-solver is not a usable reference."""
    import mpmath as mp
    mp.mp.dps = 60
    Am = mp.matrix(A.tolist()); gm = mp.matrix(g.tolist()); xm = mp.matrix(x0.tolist())
    Ainv = Am ** -1
    r = -(Ainv * gm)
    E = mp.expm(Am * dt)
    y = xm - r
    x = r + E * y
    av = r + Ainv * ((E - mp.eye(3)) * y) / dt
    return np.array([float(v) for v in x]), np.array([float(v) for v in av])


# ---------------------------------------------------------------------------------------------------------------
def test_literal_table():
    """SURVEY 8a literal table: constants the oracle derives must carry the binary32 rounding of the Fortran literals."""
    assert f32(3.141592654) == 3.1415927410125732
    assert f32(0.074) == 0.07400000095367432
    assert f32(6.346e-18) == 6.346000205740964e-18
    assert f32(13.598) == 13.597999572753906
    assert EV2K == 11604.966796875
    assert f32(0.241838e15) == 241837998604288.0
    assert f32(0.241838e15) * f32(13.598) == 3288513001696768.0
    assert COLH0 == 5.835410275968903e-11
    assert f32(1.0e-7) == 1.0000000116860974e-07 and f32(1.0e-4) == 9.999999747378752e-05
    assert float(np.sqrt(np.float32(3.0))) == 1.7320507764816284 and float(np.sqrt(np.float32(2.0))) == 1.4142135381698608
    oracle_setup(synth.make_problem(1, n=8))
    bd = O.band_data()
    assert bd["sigma_HI"][0] == f32(6.346e-18) and bd["freq_min"][0] == 3288513001696768.0
    assert bd["freq_max"][-1] == f32(0.241838e15) * f32(54.416) * 100.0


def test_rec_colion_against_numpy_restatement():
    for T in [10.0, 300.0, 8999.0, 9000.0, 1.0e4, 2.5e4, 1e5, 1e7]:
        assert relerr(O.rec_colion(T), np_rec_colion(T), 1e-300) < 5e-14, T
    r = O.rec_colion(1.0e4)  # Hui & Gnedin case B at 1e4 K ~ 2.59e-13, He0 B ~ 2.6e-13 (cgsconstants.f90:273-279)
    assert abs(r[1] / 2.59182e-13 - 1) < 1e-3 and abs(r[3] / 2.61613e-13 - 1) < 1e-3 and abs(r[6] / 1.54528e-12 - 1) < 1e-3


def _ion15(h, he, h_av=None, he_av=None, h_old=None, he_old=None):
    h_av = h if h_av is None else h_av; he_av = he if he_av is None else he_av
    h_old = h if h_old is None else h_old; he_old = he if he_old is None else he_old
    return np.array(list(h) + list(he) + list(h_av) + list(he_av) + list(h_old) + list(he_old))


@pytest.mark.parametrize("phi,dt,rtol", [((1e-12, 3e-13, 1e-14), 1e13, 1e-9), ((1e-11, 1e-12, 1e-13), 3e12, 1e-9),
                                         ((1e-14, 1e-15, 1e-17), 1e14, 1e-9),
                                         # no photons: aihe1 = n_e*colli_HeII ~ 1e-41 /s and the closed form's coefficients
                                         # (doric.f90:187-212, divisions by 2*aihe1) cancel catastrophically: the
                                         # reference's own formula is only good to ~1e-5 here, and so is any restatement
                                         ((0.0, 0.0, 0.0), 1e13, 1e-4)])
def test_doric_against_matrix_exponential(phi, dt, rtol):
    """doric's closed form (doric.f90:168-224, :267-289) must equal the matrix-exponential solution of its own ODE
    (cases where no fraction is clipped at epsilon, so the post-processing :232-258 is the identity)."""
    oracle_setup(synth.make_problem(1, n=8))
    h, he = (0.7, 0.3), (0.6, 0.3, 0.1)
    fr = (0.3, 0.6, 0.2, 0.5)
    rhe, T = 1.2e-4, 1.5e4
    out = O.doric(dt, rhe, 2e-4, _ion15(h, he), phi, fr, T)
    A, g, x0 = np_doric(dt, rhe, dict(h=h, he=he, h_old=h, he_old=he), phi, fr, T)
    x, av = expm_solution(A, g, x0, dt)
    assert 1 - x[1] - x[2] > 1e-6 and 1 - x[0] > 1e-6
    assert np.allclose([out[1], out[3], out[4]], x, rtol=rtol, atol=1e-12)
    assert np.allclose([out[6], out[8], out[9]], av, rtol=rtol, atol=1e-12)
    assert abs(out[0] + out[1] - 1) < 1e-15 and abs(out[2] + out[3] + out[4] - 1) < 1e-15


def test_doric_limits():
    oracle_setup(synth.make_problem(1, n=8))
    h, he = (0.4, 0.6), (0.5, 0.4, 0.1)
    phi, fr = (3e-17, 2e-17, 1e-17), (0.3, 0.6, 0.2, 0.5)  # rates comparable to n_e*alpha: no fraction is clipped
    o0 = O.doric(1e-3, 1e-4, 2e-4, _ion15(h, he), phi, fr, 1e4)  # dt -> 0 : old state
    assert np.allclose(o0[:5], list(h) + list(he), atol=1e-12) and np.allclose(o0[5:10], list(h) + list(he), atol=1e-12)
    big = O.doric(1e22, 1e-4, 2e-4, _ion15(h, he), phi, fr, 1e4)  # dt -> inf : equilibrium, independent of the start
    big2 = O.doric(1e22, 1e-4, 2e-4, _ion15((0.9, 0.1), (0.1, 0.1, 0.8)), phi, fr, 1e4)
    assert np.allclose(big[:5], big2[:5], rtol=1e-10, atol=1e-14)
    A, g, _ = np_doric(1e22, 1e-4, dict(h=h, he=he, h_old=h, he_old=he), phi, fr, 1e4)
    xeq = np.array([big[1], big[3], big[4]])
    assert np.all(np.abs(A @ xeq + g) <= 1e-9 * (np.abs(A) @ np.abs(xeq) + np.abs(g)))


def test_coolin_and_thermal():
    logT, *cols = O.read_cooling_table()
    oracle_setup(synth.make_problem(1, n=8))
    xh, xhe = np.array([0.3, 0.7]), np.array([0.2, 0.5, 0.3])
    for T in (12.3, 9.99e3, 1.0e4, 3.3e5, 8.7e8):
        tpos = (np.log10(T) - logT[0]) / (logT[1] - logT[0]) + 1.0
        it = min(800, max(1, int(tpos)))
        d = tpos - it
        lam = [10.0 ** c[it - 1] + (10.0 ** c[min(801, it + 1) - 1] - 10.0 ** c[it - 1]) * d for c in cols]
        ref = 2e-4 * 1e-4 * ((xh[0] * lam[0] + xh[1] * lam[1]) * (1 - ABU_HE) + (xhe[0] * lam[2] + xhe[1] * lam[3] + xhe[2] * lam[4]) * ABU_HE)
        assert abs(O.coolin(2e-4, 1e-4, xh, xhe, T) / ref - 1) < 1e-13
    # thermal: no heating -> cools; strong heating -> heats; sub-step count bounded; T<=minitemp leaves avg untouched
    ion = _ion15((0.01, 0.99), (0.01, 0.9, 0.09))
    O.set_params(False, 1e4, 1.0, 9.0, 0.0, 0.27, False, 10, 1150)
    Tend, Tav, ns = O.thermal(3e13, 2e4, 2e-4, 2e-4, ion, 0.0)
    assert Tend < 2e4 and Tend < Tav < 2e4 and 1 <= ns <= 10001
    Tend, Tav, ns = O.thermal(3e13, 1e4, 2e-4, 2e-4, ion, 1e-24)
    assert Tend > 1e4 and 1e4 < Tav < Tend
    Tend, Tav, ns = O.thermal(3e13, 0.5, 2e-4, 2e-4, ion, 0.0)
    assert ns == 0 and Tend == 0.5 and Tav == 0.0


def test_romberg_weights_and_tables():
    oracle_setup(synth.make_problem(1, n=8))
    w = O.romw()
    assert abs(w.sum() - 512.0) < 5e-4  # the b_k are binary32 (romberg.f90:53), so only ~1e-7 relative
    x = np.linspace(0.0, 1.0, 513)
    for k in range(0, 8):
        assert abs((x ** k * w).sum() / 512.0 - 1.0 / (k + 1)) < 1e-6
    info = O.sed_info()
    assert info["bb"] == (1, 33)  # T_eff=5e4 K: first band with freq_min*h/kT > 25 is 34
    thick, thin = O.table(0, 0), O.table(0, 1)
    # whole-range vs per-band quadrature of the BB photon rate agree to quadrature error (radiation_tables.f90:404)
    assert abs(thick[:, 0].sum() / 1e48 - 1) < 0.03
    assert np.all(np.diff(thick[:33], axis=1) <= 0)  # transmitted photons decrease with optical depth
    assert np.all(thin[:33, :1000] > 0)  # (bands above the limit are tabulated too, just never looked up)
    # thin table ~ -d(thick)/d(tau) at small tau: thick(0)-thick(tau) ~ tau*thin(0)
    tau = 10.0 ** (-20 + 0.012 * np.arange(2000))
    k = 1400
    assert abs((thick[0, 0] - thick[0, k + 1]) / (tau[k] * thin[0, 0]) - 1) < 1e-2
    heat = O.table(0, 2)
    assert heat.shape == (113, 2001) and np.all(heat[0] >= 0)


def test_qpl_band_limits():
    p = synth.make_problem(3, n=8, num_src=2)
    oracle_setup(p)
    assert O.sed_info()["qpl"] == (38, 47)  # 0.3 keV .. 100 nu_HeII
    t = O.table(2, 0)
    assert abs(t[37:, 0].sum() / 1e48 - 1) < 0.15


def test_photoion_rates_invariants():
    p = synth.make_problem(1, n=8)
    oracle_setup(p)
    rng = np.random.default_rng(0)
    n = 500
    lin = 10.0 ** rng.uniform(12, 22, (n, 3)); d = 10.0 ** rng.uniform(10, 21, (n, 3))
    col6 = np.empty((n, 6)); col6[:, 0::2] = lin; col6[:, 1::2] = lin + d
    vol = np.full(n, 1e66)
    out = O.photoion_rates_batch(col6, vol, [1e5, 0, 0], np.full(n, 1e-3))
    assert np.all(out[:, 4] >= out[:, 5]) and np.all(out[:, 5] >= 0)  # photons out <= photons in
    out2 = O.photoion_rates_batch(col6, vol, [2e5, 0, 0], np.full(n, 1e-3))
    assert relerr(out2, 2 * out, 1e-300) < 1e-12  # linear in NormFlux
    out3 = O.photoion_rates_batch(col6, 2 * vol, [1e5, 0, 0], np.full(n, 1e-3))
    assert relerr(out3[:, :4], out[:, :4] / 2, 1e-300) < 1e-12  # rates per volume
    # photon conservation of the cell: absorbed photons = in - out = sum over species of (rate * vol), secondary
    # ionisations excluded (isothermal tables carry no heating)
    oracle_setup(synth.make_problem(1, n=8, isothermal=True))
    o = O.photoion_rates_batch(col6, vol, [1e5, 0, 0], np.full(n, 1e-3))
    assert np.all(np.abs((o[:, 0] + o[:, 1] + o[:, 2]) * vol - (o[:, 4] - o[:, 5])) <= 1e-9 * o[:, 4])


@pytest.mark.parametrize("cfg,n,nsrc,sub,pos", [(1, 16, 1, 10, None), (1, 13, 1, 3, (2, 12, 7)), (3, 12, 3, 4, None), (2, 10, 2, 10, None)])
def test_serial_and_shell_order_sweeps_are_bitwise_identical(cfg, n, nsrc, sub, pos):
    """the max-norm shell wavefront reads the same upstream values as evolve2D's serial order."""
    p = synth.make_problem(cfg, n=n, num_src=nsrc)
    p["subboxsize"] = sub
    if pos is not None:
        p["srcpos"][0] = pos
    oracle_setup(p)
    res = []
    for order in (0, 1):
        g = oracle_grid(p)
        g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"])
        g.set_rates_to_zero()
        upd, nbox, loss, sn = g.pass_all_sources(order=order)
        res.append((upd, list(nbox), loss, g.get_rates()))
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    for a, b in zip(res[0][3], res[1][3]):
        assert np.array_equal(a, b)
    assert relerr(res[0][2], res[1][2], 1e-300) < 1e-12  # the loss is summed in a different order


def test_subbox_bookkeeping_and_latent_coverage_gap():
    """evolve_source.F90:103-144: reach left mesh/2, right mesh/2-1 (even mesh); the do-while tests dimension 3 only and
    stops as soon as last_r reaches lastpos_r, so for (mesh/2-1) mod subboxsize == 0 the left-most slab is never traced
    (SURVEY H4).  The restatement must reproduce both."""
    for n, sub, expect_cells, expect_nbox in [(16, 10, 16 ** 3, 1), (22, 10, 21 ** 3, 1), (16, 3, 16 ** 3, 3), (15, 4, 15 ** 3, 2)]:
        p = synth.make_problem(1, n=n)
        p["subboxsize"] = sub
        p["NormFlux"] = p["NormFlux"] * 1e3  # bright enough to escape every sub-box
        oracle_setup(p)
        g = oracle_grid(p)
        ion = p["xh"].copy(); ion[0] = 1e-6; ion[1] = 1 - 1e-6  # transparent box
        g.set_work_state(ion, p["xhe"], ion, p["xhe"])
        g.set_rates_to_zero()
        upd, nbox, loss, sn = g.pass_all_sources()
        assert upd == expect_cells, (n, sub, upd)
        assert list(nbox) == [expect_nbox] and sn == expect_nbox
        assert loss > 0
    # opaque, faint source: the first sub-box suffices
    p = synth.make_problem(1, n=32)
    p["subboxsize"] = 4
    p["NormFlux"] = p["NormFlux"] * 1e-9
    oracle_setup(p)
    g = oracle_grid(p)
    g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"])
    g.set_rates_to_zero()
    upd, nbox, loss, sn = g.pass_all_sources()
    assert upd < 32 ** 3 and nbox[0] < 4 and upd == (2 * 4 * nbox[0] + 1) ** 3


def test_evolve3d_invariants():
    p = synth.make_problem(1, n=16)
    oracle_setup(p)
    g = oracle_grid(p)
    before = g.state_sums(p["xh"], p["xhe"])
    st = g.evolve3d(p["dt"])
    xh, xhe, T = g.get_state()
    assert 2 <= st["niter"] <= 500 and st["conv_flag"] < st["conv_criterion"] or st["niter"] > 500
    assert np.abs(xh.sum(axis=0) - 1).max() < 1e-12 and np.abs(xhe.sum(axis=0) - 1).max() < 1e-12
    assert xh.min() >= 1e-20 and xhe.min() >= 1e-20
    after = g.state_sums(xh, xhe)
    assert abs((after[0] + after[1]) / (before[0] + before[1]) - 1) < 1e-12  # H nuclei conserved
    c = 8  # source cell index (srcpos = n//2 = 8, 1-based) -> [7]
    assert xh[1, 7, 7, 7] > 0.99 and xh[1, 0, 0, 0] < 1e-3  # ionized at the source, neutral in the far corner
    assert np.array_equal(T[2], T[0]) and T.max() > 1.0e4  # set_final_temperature_point; photo-heated gas
    # photon budget: new ionizations cannot exceed the photons emitted in dt (1e55 /s)
    ions = (before[0] - after[0]) + (before[2] - after[2]) + (after[4] - before[4])
    assert 0.3 < ions / (1e55 * p["dt"]) < 1.05


def test_mrgrnk_is_a_stable_rank():
    rng = np.random.default_rng(2)
    x = rng.integers(0, 50, 1000).astype(np.float32)  # many ties
    assert np.array_equal(O.mrgrnk(x), np.argsort(x, kind="stable") + 1)


def test_arithmetic_noise_floor(tmp_path):
    """The same restatement compiled with FMA contraction differs from itself by ~1e-10 absolute in the fractions: the
    floor below which a relative criterion on small fractions is meaningless (tests/common.py FRAC_ATOL)."""
    res = []
    for libname in ("libc2ray_oracle.so", "libc2ray_oracle_fma.so"):
        out = tmp_path / (libname + ".npz")
        env = dict(os.environ, C2RAY_ORACLE_LIB=libname)
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "oracle_run.py"), "chem", str(out)], env=env)
        res.append(np.load(out))
    a, b = res
    worst_abs = 0.0
    for k in ("ion_0", "ion_1"):
        assert np.array_equal(a["nit_" + k[-1]], b["nit_" + k[-1]])
        assert frac_err(a[k][:, :10], b[k][:, :10]) < 1
        worst_abs = max(worst_abs, np.abs(a[k][:, :10] - b[k][:, :10]).max())
        small = a[k][:, :10] < 1e-6
        assert relerr(a[k][:, :10][small], b[k][:, :10][small], 1e-300) > 1e-8  # a pure relative test would fail here
    assert 1e-13 < worst_abs < 2e-10


def test_golden_fixtures():
    """tests/golden/*.npz were written by tools/make_golden.py from this oracle (there is no reference output to pin
    against): a regression pin for the restatement and the fixture the GPU tests reuse."""
    import make_golden
    g = np.load(os.path.join(ROOT, "tests", "golden", "hotpath_small.npz"))
    now = make_golden.compute()
    for k in g.files:
        if g[k].dtype.kind in "iu":
            assert np.array_equal(g[k], now[k]), k
        elif k.startswith("frac_"):
            assert frac_err(now[k], g[k]) < 1e-3, k   # same binary: far below the tolerance
        else:
            assert relerr(now[k], g[k], 1e-300) < 1e-12, k
