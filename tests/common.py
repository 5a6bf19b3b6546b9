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


def load_oracle_variant(libname="libc2ray_oracle_fma.so"):
    """A second, independent instance of the oracle module bound to another build of the same source (default: the one
    compiled with -mfma -ffp-contract=fast).  Used to measure how far the reference's own arithmetic determines a result:
    the two CPU builds differ only in rounding."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(O.__file__)), "oracle.py")
    spec = importlib.util.spec_from_file_location("oracle_variant_" + libname.replace(".", "_"), path)
    mod = importlib.util.module_from_spec(spec)
    old = os.environ.get("C2RAY_ORACLE_LIB")
    os.environ["C2RAY_ORACLE_LIB"] = libname
    try:
        O.build()
        spec.loader.exec_module(mod)
        mod.lib()
    finally:
        if old is None:
            os.environ.pop("C2RAY_ORACLE_LIB", None)
        else:
            os.environ["C2RAY_ORACLE_LIB"] = old
    return mod


VARIANT_LIBS = ("libc2ray_oracle_fma.so", "libc2ray_oracle_assoc.so")


def load_oracle_variants():
    """Both calibration builds of the oracle source: FMA contraction allowed, and the reassociating -O3 build that stands for
    the reference's own production compilers (files_for_3D/Makefile:63-64 ifort -O3 -ipo, :111 pgf90 -O3 -fast)."""
    return [load_oracle_variant(n) for n in VARIANT_LIBS]


def setup_variant(mod, p):
    """oracle_setup + oracle_grid for a variant module (its own library-global state)."""
    iso = p["isothermal"]
    mod.rad_ini(p["T_eff"], p["S_star"], pl=p.get("pl"), qpl=p.get("qpl"), isothermal=iso)
    mod.set_params(iso, p["temper_val"], p["clumping"], p["zred"], p["H0"], p["Omega0"], p["cosmological"],
                   p.get("subboxsize", 10), p.get("max_subbox", 1150))
    g = mod.Grid(p["mesh"], p["dr"], p["vol"])
    g.set_state(p["ndens"], p["xh"], p["xhe"], p["temperature_grid"])
    g.set_sources(p["srcpos"], p["NormFlux"], p.get("NormFluxPL"), p.get("NormFluxQPL"))
    return g


# ---- comparison calibrated on the reference's own reproducibility (full-size parity, independent restatements) --------
EDGES = [1e-16, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-4, 1e-2, 1.0]
NOISE_FACTOR = 10.0
COMPLEMENT_ULPS = 2 * 2.220446049250313e-16   # fractions are stored as complements (doric.f90:222-224): two ulps of 1.0


def temperature_bound(e_cpu):
    """Bound on the relative difference of float32 temperatures (mat_ini_test.F90:31), given the largest difference between
    the oracle's own builds: one float ulp when those agree exactly, one ulp more when they are themselves an ulp apart
    somewhere (their double-precision values straddle a rounding boundary: a third value can be one step further), ten
    times their difference beyond that."""
    if e_cpu == 0.0:
        return 1.3e-7
    if e_cpu <= 1.3e-7:
        return 2.5e-7
    return NOISE_FACTOR * e_cpu


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    nz = ref != 0.0
    e = np.zeros(ref.shape)
    e[nz] = np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])
    return e, nz


def hist_block(e, nz, ref):
    rel = e[nz]
    bad = rel > 1e-8
    out = {"max_rel": float(rel.max()) if rel.size else 0.0, "exceed_1e-8": int(bad.sum()),
           "hist_counts": [int(x) for x in np.histogram(rel, bins=[0.0] + EDGES + [np.inf])[0]]}
    if bad.any():
        r = np.abs(ref[nz][bad])
        out["exceeders_abs_ref"] = {"max": float(r.max()), "median": float(np.median(r))}
    return out


def calibrated_compare(got, ref, alt, atol=0.0, rtol=1e-8):
    """Pure-relative comparison calibrated on the reference's own reproducibility (tests/test_gpu_fullsize.py docstring).
    got: implementation under test, ref: oracle, alt: the same field from the oracle's other build(s) -- one array or a
    list of arrays (FMA contraction allowed; reassociating -O3), whose cell-wise largest deviation from ref is the noise;
    bound per decade of |ref|/max|ref| = max(rtol, NOISE_FACTOR x largest |alt-ref|/|ref| of that decade and its two
    neighbours); atol: absolute term (fractions only: COMPLEMENT_ULPS).  Returns (record, ok)."""
    ref = np.asarray(ref, dtype=np.float64).ravel()
    e_gpu, nz = rel_err(got, ref)
    alts = list(alt) if isinstance(alt, (list, tuple)) else [alt]
    e_cpu = np.zeros(ref.shape)
    alt_nonzero = np.zeros(ref.shape, dtype=np.int64)
    for a in alts:
        e_cpu = np.maximum(e_cpu, rel_err(a, ref)[0])
        alt_nonzero = np.maximum(alt_nonzero, ((np.asarray(a).ravel() != 0.0) != nz).astype(np.int64))
    # exact zeros: untraced cells, and cells whose own column is absorbed by the rounding of the incoming one
    # ((in + c) - in == 0, radiation_photoionrates.f90:167-169) -- a knife edge the two CPU builds also disagree on
    rec = {"cells": int(ref.size), "ref_nonzero": int(nz.sum()),
           "zero_pattern_mismatch": int(((np.asarray(got).ravel() != 0.0) != nz).sum()),
           "zero_pattern_mismatch_cpu_cpu": int(alt_nonzero.sum()), "cpu_builds_compared": len(alts),
           "peak_abs_ref": float(np.abs(ref).max()) if ref.size else 0.0, "hist_edges": EDGES,
           "gpu_vs_oracle": hist_block(e_gpu, nz, ref), "oracle_builds_vs_oracle": hist_block(e_cpu, nz, ref)}
    ok = rec["zero_pattern_mismatch"] <= NOISE_FACTOR * rec["zero_pattern_mismatch_cpu_cpu"]
    if nz.any():
        peak = np.abs(ref).max()
        dec = np.full(ref.shape, 0, dtype=np.int64)
        dec[nz] = np.clip(np.floor(-np.log10(np.abs(ref[nz]) / peak)), 0, 330).astype(np.int64)  # decades below the peak
        nd = int(dec[nz].max()) + 1
        env = np.zeros(nd + 2)
        np.maximum.at(env, dec[nz] + 1, e_cpu[nz])
        env3 = np.maximum(np.maximum(env[:-2], env[1:-1]), env[2:])          # this decade and its two neighbours
        bound = np.maximum(rtol, NOISE_FACTOR * env3)
        viol = nz & (e_gpu * np.abs(ref) > bound[np.minimum(dec, nd - 1)] * np.abs(ref) + atol)
        gmax = np.zeros(nd)
        np.maximum.at(gmax, dec[nz], e_gpu[nz])
        cnt = np.bincount(dec[nz], minlength=nd)
        rec["by_decade_below_peak"] = [{"decade": d, "cells": int(cnt[d]), "gpu_max": float(gmax[d]), "cpu_cpu_max": float(env[d + 1]),
                                        "bound": float(bound[d])} for d in range(nd) if cnt[d]]
        rec["violations"] = int(viol.sum())
        rec["failed_decades"] = [(d["decade"], d["cells"], d["gpu_max"], d["cpu_cpu_max"], d["bound"])
                                 for d in rec["by_decade_below_peak"] if d["gpu_max"] > d["bound"]]
        # first decade (counted from the peak) in which the reference stops agreeing with itself to 1e-9
        noisy = [d for d in range(nd) if env[d + 1] > 1e-9]
        rec["well_conditioned_down_to_decade"] = int(noisy[0]) if noisy else nd
        ok = ok and rec["violations"] == 0
    return rec, ok
