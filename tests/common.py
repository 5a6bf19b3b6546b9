"""Shared helpers of the parity tests: the oracle is the checker, the C ABI library is the thing under test."""
import numpy as np

import c2ray_b200
from oracle import oracle as O

_tables_key = None


def oracle_setup(p, isothermal=None):
    """(Re)initialise the oracle's module state for problem p; returns the table dict for C2Ray.upload_tables."""
    global _tables_key
    iso = p["isothermal"] if isothermal is None else isothermal
    qpl = p.get("qpl")
    pl = p.get("pl")
    key = (p["T_eff"], p["S_star"], None if qpl is None else tuple(sorted(qpl.items())),
           None if pl is None else tuple(sorted(pl.items())), iso)
    if key != _tables_key:
        O.rad_ini(p["T_eff"], p["S_star"], pl=pl, qpl=qpl, isothermal=iso)
        _tables_key = key
    O.set_params(iso, p["temper_val"], p["clumping"], p["zred"], p["H0"], p["Omega0"], p["cosmological"],
                 p.get("subboxsize", 10), p.get("max_subbox", 1150))
    info = O.sed_info()
    tables = {0: tuple(O.table(0, k) if (k < 2 or not iso) else None for k in range(4)) + (info["bb"][0], info["bb"][1], p["S_star"])}
    if pl is not None:
        tables[1] = tuple(O.table(1, k) if (k < 2 or not iso) else None for k in range(4)) + (info["pl"][0], info["pl"][1], pl["S_star"])
    if qpl is not None:
        tables[2] = tuple(O.table(2, k) if (k < 2 or not iso) else None for k in range(4)) + (info["qpl"][0], info["qpl"][1], qpl["S_star"])
    return tables


def oracle_grid(p):
    g = O.Grid(p["mesh"], p["dr"], p["vol"])
    g.set_state(p["ndens"], p["xh"], p["xhe"], p["temperature_grid"])
    g.set_sources(p["srcpos"], p["NormFlux"], p.get("NormFluxPL"), p.get("NormFluxQPL"))
    return g


def relerr(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def partially_ionized_state(p, seed=1):
    """A non-trivial (xh_av, xhe_av) so that sweeps see ionized and neutral gas."""
    rng = np.random.default_rng(seed)
    n = p["mesh"][0]
    shape = (n, n, n)
    x1 = 10.0 ** rng.uniform(-6, 0, shape) * 0.999
    xh = np.stack([1.0 - x1, x1])
    a = 10.0 ** rng.uniform(-6, 0, shape) * 0.6
    b = 10.0 ** rng.uniform(-8, 0, shape) * 0.39
    xhe = np.stack([1.0 - a - b, a, b])
    return xh, xhe


# Ionization fractions: 1e-8 relative (BASELINE north_star) plus an absolute floor.  doric's closed-form solution
# (doric.f90:222-224, :285-287) forms small fractions as differences of O(1) terms, so a fraction x carries an absolute
# rounding noise of order 1e-11..1e-10 whatever its size: the CPU oracle compiled with and without FMA contraction
# differs from itself by up to 8.6e-11 (tests/test_oracle_cpu.py::test_arithmetic_noise_floor measures this).
FRAC_RTOL, FRAC_ATOL = 1e-8, 2e-10


def frac_err(a, b):
    """max |a-b| / (FRAC_RTOL*|b| + FRAC_ATOL) ; parity holds when < 1"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (FRAC_RTOL * np.abs(b) + FRAC_ATOL)))
