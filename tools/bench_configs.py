"""All five BASELINE.json configs on one GPU, one JSON line each (the headline bench.py covers configs[1] only).

  python tools/bench_configs.py [--out gpurun_out/configs.jsonl] [--only 1,3] [--chem-n 512]

configs[0]  128^3 uniform box, one BB source 5e4 K, subboxsize 10            full evolve3D time step
configs[1]  128^3 Test-4 style, 16 BB sources 1e5 K                            full evolve3D time step (= bench.py)
configs[2]  256^3, 1000 sources, BB + QPL on the 50 brightest, subboxsize 10   two global iterations (RT pass + pass)
configs[3]  512^3, this GPU's share (1250 of 10^4 sources) of the 8-GPU run    two global iterations
configs[4]  chemistry only, one 1024x1024x128 slab of the 1024^3 box per GPU   global pass, thermal and isothermal
Device times are CUDA-event times on the context's stream; every number is the mean of the timed repetitions after one
untimed repetition.  Roofline fractions use MEASURED_PEAKS.json (6553 GB/s) and the algorithmic bytes of SURVEY 8d."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import c2ray_b200

PEAK = 6553.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def time_step(cfg, n, nsrc, steps=2):
    p = c2ray_b200.synth.make_problem(cfg, n=n, num_src=nsrc, isothermal=False)
    c = c2ray_b200.from_problem(p, device=0)
    c.snapshot_state()
    c.evolve3D(0.0, p["dt"], 0)
    out = []
    for _ in range(steps):
        c.restore_state()
        c.timer_start()
        s = c.evolve3D(0.0, p["dt"], 0)
        ms = c.timer_stop()
        out.append((ms, s))
    c.close()
    ms = np.mean([o[0] for o in out]); s = out[-1][1]
    upd = s["rt_updates"]
    return {"mesh": n, "sources": nsrc, "what": "full evolve3D time step", "niter": s["niter"], "s_per_timestep": ms * 1e-3,
            "rt_updates_per_step": int(upd), "updates_per_s": upd / (ms * 1e-3),
            "sweep_updates_per_s": upd / (s["ms_sweep"] * 1e-3), "chem_cells_per_s": s["chem_cells"] / (s["ms_chem"] * 1e-3),
            "ms_sweep": s["ms_sweep"], "ms_chem": s["ms_chem"], "photcons": s["photcons"],
            "sweep_hbm_frac": 104 * upd / (s["ms_sweep"] * 1e-3) / 1e9 / PEAK}


def iterations(cfg, n, nsrc, iters=3):
    t0 = time.time()
    p = c2ray_b200.synth.make_problem(cfg, n=n, num_src=nsrc)
    c = c2ray_b200.from_problem(p, device=0)
    c.begin_step()
    rows = []
    for it in range(iters):
        c.set_rates_to_zero()
        c.timer_start(); upd = c.pass_all_sources(it + 1, p["dt"]); ms_s = c.timer_stop()
        c.timer_start(); cf = c.global_pass(p["dt"]); ms_c = c.timer_stop()
        rows.append((upd, ms_s, ms_c, cf))
    c.close()
    rows = rows[1:]  # the first iteration allocates the sweep scratch
    upd = np.mean([r[0] for r in rows]); ms_s = np.mean([r[1] for r in rows]); ms_c = np.mean([r[2] for r in rows])
    return {"mesh": n, "sources": nsrc, "what": f"global iterations 2..{iters} from the neutral start state (RT pass + global pass)",
            "rt_updates_per_pass": int(upd), "coverage": upd / nsrc / n ** 3, "ms_sweep": ms_s, "ms_chem": ms_c,
            "sweep_updates_per_s": upd / (ms_s * 1e-3), "chem_cells_per_s": n ** 3 / (ms_c * 1e-3),
            "sweep_hbm_frac": 104 * upd / (ms_s * 1e-3) / 1e9 / PEAK, "chem_hbm_frac": 224 * n ** 3 / (ms_c * 1e-3) / 1e9 / PEAK,
            "conv_flag": int(rows[-1][3]), "setup_s": time.time() - t0}


def chemistry(mesh, iso, reps=3):
    ncell = int(np.prod(mesh))
    q = c2ray_b200.synth.make_chemistry_problem(ncell, isothermal=iso)
    par = c2ray_b200.C2RayParameters(isothermal=iso, H0=q["H0"], Omega0=q["Omega0"])
    c = c2ray_b200.C2Ray(mesh, par, device=0)
    c.setup_cool()
    c.set_geometry([1e22] * 3, 1e66, q["zred"])
    c.set_state(q["ndens"], q["xh"], q["xhe"], q["temperature_grid"])
    c.snapshot_state()
    c.set_rates(q["phih"], q["phihe"], q["phiheat"])
    c.bench_global_pass(q["dt"], 1)
    ms, cf = c.bench_global_pass(q["dt"], reps)
    c.close()
    B = 200 if iso else 224
    return {"mesh": list(mesh), "cells": ncell, "what": f"global pass, {'isothermal' if iso else 'thermal'}, config-5 inputs",
            "ms_per_pass": ms, "chem_cells_per_s": ncell / (ms * 1e-3), "bytes_per_cell": B,
            "achieved_gbs": B * ncell / (ms * 1e-3) / 1e9, "chem_hbm_frac": B * ncell / (ms * 1e-3) / 1e9 / PEAK, "conv_flag": int(cf)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="0,1,2,3,4")
    ap.add_argument("--chem-n", type=int, default=1024)
    a = ap.parse_args()
    only = {int(x) for x in a.only.split(",")}
    lines = []

    def emit(idx, d):
        d = {"config": idx, **d}
        lines.append(d)
        print(json.dumps(d), flush=True)

    if 0 in only:
        emit(0, time_step(1, 128, 1))
    if 1 in only:
        emit(1, time_step(2, 128, 16))
    if 2 in only:
        emit(2, iterations(3, 256, 1000))
    if 3 in only:
        emit(3, iterations(4, 512, 1250))
    if 4 in only:
        n = a.chem_n
        emit(4, chemistry([n, n, max(1, n // 8)], False))
        emit(4, chemistry([n, n, max(1, n // 8)], True))
    if a.out:
        with open(a.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
