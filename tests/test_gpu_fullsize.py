"""Full-size GPU-vs-oracle parity on the BASELINE configs.

configs[0] (128^3, 1 BB source, subboxsize 10) and configs[1] (128^3 Test-4 style, 16 BB sources, subboxsize = mesh):
one complete evolve3D time step through the product entry point against the oracle's evolve3D (evolve.F90:120-229) --
every integer of the step exactly (niter, conv_flag after every global iteration, RT updates, sum_nbox), the final
fractions / temperature / rate grids cell by cell.  configs[2] (256^3, BB + QPL sources, subboxsize 10) on an
8-source subset of the 1000-source list, three global iterations through the stepwise calls, compared
after every pass (evolve.F90:154-222, evolve_point.F90:406-424).

Every comparison also records PURE-RELATIVE statistics |got-ref|/|ref| (no absolute floor, ref != 0) per array:
histogram by decade, how many cells exceed 1e-8 and how large those cells' values are.  They are printed and written
to gpurun_out/parity_fullsize_<name>.json; the committed copies live in profiles/.

The rate grids are asserted pure-relative (1e-8 wherever ref != 0).  Fractions are asserted with tests/common.py's
floor (|dx| <= 1e-8 x + 2e-10), and the pure-relative offenders are bounded: they must all be small fractions
(x < 1e-2), i.e. inside doric's cancellation noise.
"""
import json
import os
import time

import numpy as np
import pytest

import c2ray_b200
from oracle import oracle as O
from common import oracle_setup, oracle_grid, frac_err, FRAC_RTOL, FRAC_ATOL

pytestmark = pytest.mark.gpu
synth = c2ray_b200.synth
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EDGES = [1e-16, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-4, 1e-2, 1.0]


def rel_stats(got, ref):
    """Pure-relative comparison of two arrays where ref != 0 (and a count of cells where exactly one side is 0)."""
    got = np.asarray(got, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    nz = ref != 0.0
    rel = np.abs(got[nz] - ref[nz]) / np.abs(ref[nz])
    hist = np.histogram(rel, bins=[0.0] + EDGES + [np.inf])[0]
    bad = rel > 1e-8
    out = {"cells": int(ref.size), "ref_nonzero": int(nz.sum()), "zero_pattern_mismatch": int(((got != 0.0) != nz).sum()),
           "max_rel": float(rel.max()) if rel.size else 0.0, "exceed_1e-8": int(bad.sum()),
           "hist_edges": EDGES, "hist_counts": [int(x) for x in hist]}
    if bad.any():
        r = np.abs(ref[nz][bad])
        out["exceeders"] = {"max_abs_ref": float(r.max()), "median_abs_ref": float(np.median(r)),
                            "max_abs_diff": float(np.abs(got[nz][bad] - ref[nz][bad]).max())}
    return out


def dump(name, record):
    print(f"\n[parity {name}] " + json.dumps(record)[:4000])
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, f"parity_fullsize_{name}.json"), "w") as f:
            json.dump(record, f, indent=1)
    except OSError:
        pass


def compare_state(tag, rec, got, ref, iso):
    """got/ref: (xh, xhe, T) or (xh_av, xhe_av, xh_int, xhe_int).  Returns the worst floor-normalised fraction error."""
    worst = 0.0
    for name, a, b in zip(tag, got, ref):
        if a.dtype == np.float32 or name.startswith("T"):
            if iso:
                continue
            rec[name] = rel_stats(a.astype(np.float64), b.astype(np.float64))
            continue
        for comp in range(a.shape[0]):
            s = rel_stats(a[comp], b[comp])
            s["floor_err"] = float(np.max(np.abs(a[comp] - b[comp]) / (FRAC_RTOL * np.abs(b[comp]) + FRAC_ATOL)))
            rec[f"{name}[{comp}]"] = s
            worst = max(worst, s["floor_err"])
    return worst


def assert_fraction_offenders_small(rec):
    for k, s in rec.items():
        if isinstance(s, dict) and "floor_err" in s and s["exceed_1e-8"]:
            assert s["exceeders"]["max_abs_ref"] < 1e-2, (k, s["exceeders"])


def full_step(cfg_index, p, name):
    nthreads = O.num_threads()
    tables = oracle_setup(p)
    c = c2ray_b200.from_problem(p, tables=tables)
    t0 = time.perf_counter()
    sg = c.evolve3D(0.0, p["dt"], 0)
    t_gpu = time.perf_counter() - t0
    xh, xhe, T = c.get_state()
    rates_g = c.get_rates()
    g = oracle_grid(p)
    t0 = time.perf_counter()
    so = g.evolve3d(p["dt"], nthreads=nthreads, order=2)
    t_cpu = time.perf_counter() - t0
    xh_o, xhe_o, T_o = g.get_state()
    rates_o = g.get_rates()
    rec = {"config": f"BASELINE configs[{cfg_index}]", "mesh": int(p["mesh"][0]), "sources": int(len(p["NormFlux"])),
           "niter_gpu": int(sg["niter"]), "niter_oracle": int(so["niter"]), "conv_hist_gpu": [int(x) for x in sg["conv_hist"]],
           "conv_hist_oracle": [int(x) for x in so["conv_hist"]], "rt_updates_gpu": int(sg["rt_updates"]),
           "rt_updates_oracle": int(so["rt_updates"]), "sum_nbox_gpu": int(sg["sum_nbox_all"]), "sum_nbox_oracle": int(so["sum_nbox"]),
           "seconds_gpu": t_gpu, "seconds_oracle": t_cpu, "oracle_threads": nthreads}
    worst = compare_state(("xh", "xhe", "T"), rec, (xh, xhe, T), (xh_o, xhe_o, T_o), p["isothermal"])
    for nm, a, b in zip(("phih", "phihe", "phiheat"), rates_g, rates_o):
        if a.ndim == 4:
            for comp in range(a.shape[0]):
                rec[f"{nm}[{comp}]"] = rel_stats(a[comp], b[comp])
        else:
            rec[nm] = rel_stats(a, b)
    dump(name, rec)
    c.close()
    # integers: exact
    assert sg["niter"] == so["niter"]
    assert list(sg["conv_hist"]) == list(so["conv_hist"])
    assert sg["rt_updates"] == so["rt_updates"]
    assert sg["sum_nbox_all"] == so["sum_nbox"]
    # rate grids of the last pass: pure relative wherever the reference value is non-zero, same zero pattern
    for nm in ("phih", "phihe[0]", "phihe[1]", "phiheat"):
        assert rec[nm]["zero_pattern_mismatch"] == 0, (nm, rec[nm])
        assert rec[nm]["max_rel"] < 1e-8, (nm, rec[nm])
    # fractions: 1e-8 relative + the noise floor; temperature: one float32 ulp
    assert worst < 1, worst
    assert_fraction_offenders_small(rec)
    assert rec["T"]["max_rel"] < 1.3e-7
    return rec


def test_config0_full_step():
    """128^3 uniform box, one 5e4 K black-body source of 1e55 photons/s, subboxsize 10, non-isothermal, dt = 5 Myr."""
    p = synth.make_problem(1, n=128)
    full_step(0, p, "config0_128_1src")


def test_config1_full_step():
    """128^3 Test-4-style lognormal box, 16 black-body sources of 1e5 K, subboxsize = mesh, non-isothermal."""
    p = synth.make_problem(2, n=128)
    full_step(1, p, "config1_128_16src")


def test_config2_subset_three_iterations():
    """256^3 lognormal box; sources 1, 2, 3 (BB + QPL, the brightest), 61, 301, 701, 901, 1000 (BB only) of the
    1000-source list; three global iterations through pass_all_sources / global_pass, compared after every pass."""
    full = synth.make_problem(3, n=256)
    pick = np.array([0, 1, 2, 60, 300, 700, 900, 999])
    p = dict(full)
    p["srcpos"] = np.ascontiguousarray(full["srcpos"][pick])
    p["NormFlux"] = np.ascontiguousarray(full["NormFlux"][pick])
    p["NormFluxQPL"] = np.ascontiguousarray(full["NormFluxQPL"][pick])
    nthreads = O.num_threads()
    tables = oracle_setup(p)
    c = c2ray_b200.from_problem(p, tables=tables)
    g = oracle_grid(p)
    g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"])
    c.begin_step()
    rec = {"config": "BASELINE configs[2], 8-source subset", "mesh": 256, "sources": [int(x) + 1 for x in pick],
           "oracle_threads": nthreads, "iterations": []}
    worst = 0.0
    t0 = time.perf_counter()
    for it in range(1, 4):
        g.set_rates_to_zero()
        c.set_rates_to_zero()
        upd_o, nbox_o, loss_o, sum_nbox_o = g.pass_all_sources(nthreads=nthreads, order=2)
        upd_g = c.pass_all_sources(it, p["dt"])
        r = {"iteration": it, "rt_updates_gpu": int(upd_g), "rt_updates_oracle": int(upd_o), "nbox_oracle": [int(x) for x in nbox_o]}
        for nm, a, b in zip(("phih", "phihe", "phiheat"), c.get_rates(), g.get_rates()):
            if a.ndim == 4:
                for comp in range(a.shape[0]):
                    r[f"{nm}[{comp}]"] = rel_stats(a[comp], b[comp])
            else:
                r[nm] = rel_stats(a, b)
        cf_o = g.global_pass(p["dt"], nthreads=nthreads)
        cf_g = c.global_pass(p["dt"])
        r["conv_flag_gpu"], r["conv_flag_oracle"] = int(cf_g), int(cf_o)
        w = compare_state(("xh_av", "xhe_av", "xh_intermed", "xhe_intermed"), r, c.get_work_state(), g.get_work_state(), False)
        Tg, To = c.get_state()[2], g.get_state()[2]
        r["T(0:1)"] = rel_stats(Tg[:2].astype(np.float64), To[:2].astype(np.float64))
        worst = max(worst, w)
        rec["iterations"].append(r)
    rec["seconds_total"] = time.perf_counter() - t0
    dump("config2_256_8src", rec)
    c.close()
    for r in rec["iterations"]:
        assert r["rt_updates_gpu"] == r["rt_updates_oracle"], r["iteration"]
        assert r["conv_flag_gpu"] == r["conv_flag_oracle"], r["iteration"]
        for nm in ("phih", "phihe[0]", "phihe[1]", "phiheat"):
            assert r[nm]["zero_pattern_mismatch"] == 0, (r["iteration"], nm, r[nm])
            assert r[nm]["max_rel"] < 1e-8, (r["iteration"], nm, r[nm])
        assert_fraction_offenders_small(r)
        assert r["T(0:1)"]["max_rel"] < 1.3e-7
    assert worst < 1, worst
