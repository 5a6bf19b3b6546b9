#!/usr/bin/env python
"""bench.py -- headline benchmark of the C2-Ray H+He hot path on B200 (see DESIGN.md "Measurement").

Metric: source x cell RT updates/s (BASELINE.json), whole job, over full evolve3D time steps (ray-tracing sweeps of all
sources + rate-grid reduction + global chemistry passes, iterated to convergence).
Workload (N=1): BASELINE configs[1] -- Test-4 style 128^3 lognormal box, 16 black-body sources (T_eff=1e5 K,
subboxsize=mesh), non-isothermal.  For N>1 every rank gets 16 sources of the same box (weak scaling); per iteration the
rate grids are reduce-scattered, every rank runs the global pass on its N^3/npr cells and the fractions the next sweep
reads are all-gathered (the reference: allreduce + replicated pass; same results, see tools/multi_gpu_check.py).
A step = one evolve3D(time,dt) from the same start state (device snapshot restored inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  torchrun ... bench.py --gpus N ...        (one rank per GPU)

--impl reference times the CPU restatement of the reference (oracle/, all host threads) on a bounded sample of the same
workload: one global iteration (RT pass over the 16 sources + global chemistry pass) per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rt_source_cell_updates_per_s"
UNIT = "updates/s"
SRC_PER_GPU = 16
MESH = 128
BYTES_PER_UPDATE = 104  # 5 FP64 state reads + read-modify-write of 4 rate grids (thermal)
# DRAM traffic of k_sweep_shell from the committed `ncu --set full` capture (profiles/r1_ncu_full_kernels_final_v4_128.csv,
# first launch = one stream group (8 sources) at shell radius 56: dram__bytes_read.sum + dram__bytes_write.sum =
# 119.1 + 21.8 MB for 8 x 75,266 = 602,128 updates)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 140.9e6
NCU_TRAFFIC_BYTES_PER_UPDATE = 234.0


def workload(n_gpus, mesh):
    import c2ray_b200
    p = c2ray_b200.synth.make_problem(2, n=mesh, num_src=SRC_PER_GPU * n_gpus, isothermal=False)
    if n_gpus > 1:  # keep every source as bright as in the 16-source box
        p["NormFlux"] = p["NormFlux"] * n_gpus
    return p


def config_dict(p, n_gpus):
    return {"workload": f"BASELINE configs[1]: Test-4-style {p['mesh'][0]}^3 lognormal box (sigma=1, seed 4), "
                        f"{len(p['NormFlux'])} BB sources T_eff=1e5 K, subboxsize=mesh, non-isothermal, dt=0.05 Myr, "
                        "one full evolve3D time step per step",
            "mesh": int(p["mesh"][0]), "sources": int(len(p["NormFlux"])), "sources_per_gpu": SRC_PER_GPU,
            "parallelism": ("one GPU" if n_gpus == 1 else
                            f"sources round-robin over {n_gpus} GPUs (do_grid_static); per iteration one reduce-scatter of the "
                            "rate grids, the global pass on N^3/npr cells per rank, one all-gather of the fractions the next "
                            "sweep reads"),
            "l2_policy": "per-step working set (state + rate grids + snapshot, >400 MB) exceeds the 126 MB L2"}


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled during the timed region."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_iteration(p, nthreads, sources=None):
    """One global iteration of the oracle (reference restatement) on host threads: RT pass + global pass.
    Returns (rt_updates, seconds_rt, seconds_chem)."""
    from oracle import oracle as O
    O.rad_ini(p["T_eff"], p["S_star"], qpl=p.get("qpl"), isothermal=p["isothermal"])
    O.set_params(p["isothermal"], p["temper_val"], p["clumping"], p["zred"], p["H0"], p["Omega0"], p["cosmological"],
                 p["subboxsize"], p["max_subbox"])
    g = O.Grid(p["mesh"], p["dr"], p["vol"])
    g.set_state(p["ndens"], p["xh"], p["xhe"], p["temperature_grid"])
    ns = len(p["NormFlux"]) if sources is None else sources
    g.set_sources(p["srcpos"][:ns], p["NormFlux"][:ns], None, None if p.get("NormFluxQPL") is None else p["NormFluxQPL"][:ns])
    g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"])
    g.set_rates_to_zero()
    return g, ns


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    n_gpus = int(os.environ.get("WORLD_SIZE", args.gpus))
    p = workload(n_gpus, args.mesh)   # the same config as the GPU arm at this N
    nthreads = O.num_threads()
    # bounded sample: at most 32 sources per step (each source costs ~2.1 M updates ~ 0.25 core-seconds x 8)
    g, ns = cpu_reference_iteration(p, nthreads, sources=min(len(p["NormFlux"]), 32))
    times, updates = [], 0
    for step in range(args.warmup + args.steps):
        g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"])
        g.set_state(p["ndens"], p["xh"], p["xhe"], p["temperature_grid"])
        g.set_rates_to_zero()
        t0 = time.perf_counter()
        upd, _, _, _ = g.pass_all_sources(nthreads=nthreads, order=0)
        g.global_pass(p["dt"], nthreads=nthreads)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt); updates += upd
    total = sum(times)
    value = updates / total
    sample = (f"per step: one global iteration (RT pass over {ns} of {len(p['NormFlux'])} sources + global chemistry pass) of the {args.mesh}^3 "
              f"workload from the neutral start state, C++ restatement of the reference, g++ -O2 -fopenmp")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(p, n_gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import c2ray_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p = workload(n_gpus, args.mesh)
    c = c2ray_b200.from_problem(p, device=local)
    if world > 1:
        uid = [c2ray_b200.C2Ray.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        c.comm_init(uid[0], rank, world)
    c.snapshot_state()
    N3 = c.N3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: value ---------------------------------------------------------------------------------
    stats = []
    for step in range(args.warmup):
        c.restore_state()
        c.evolve3D(0.0, p["dt"], 0)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = c.launch_count()
    c.timer_start()
    for step in range(args.steps):
        c.restore_state()
        stats.append(c.evolve3D(0.0, p["dt"], 0))
    ms = c.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = c.launch_count() - l0
    upd_local = sum(s["rt_updates"] for s in stats)
    ms_sweep = sum(s["ms_sweep"] for s in stats)
    ms_chem = sum(s["ms_chem"] for s in stats)
    ms_ar = sum(s["ms_allreduce"] for s in stats)

    # ---- end-to-end arm: host buffers through the Fortran-facing entry point ------------------------------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_ndens, h_xh0, h_xhe0, h_T0 = pin(p["ndens"]), pin(p["xh"]), pin(p["xhe"]), pin(p["temperature_grid"])
    h_xh, h_xhe, h_T = torch.empty_like(h_xh0).pin_memory(), torch.empty_like(h_xhe0).pin_memory(), torch.empty_like(h_T0).pin_memory()
    e2e_steps = max(1, min(args.steps, 2))
    barrier()
    t0 = time.perf_counter()
    upd_e2e = 0
    for step in range(e2e_steps):
        h_xh.copy_(h_xh0); h_xhe.copy_(h_xhe0); h_T.copy_(h_T0)
        s = c.evolve3D_host(0.0, p["dt"], 0, h_ndens.numpy(), h_xh.numpy(), h_xhe.numpy(), h_T.numpy())
        upd_e2e += s["rt_updates"]
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    h2d = N3 * (8 + 16 + 24 + 12)
    d2h = N3 * (16 + 24 + 12)

    # ---- reduce over ranks: max time, summed units --------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms, t_e2e, ms_sweep, ms_chem, ms_ar], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e, ms_sweep_max, ms_chem_max, ms_ar_max = t.tolist()
        u = torch.tensor([upd_local, upd_e2e, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        upd_total, upd_e2e_total, launches_total = u.tolist()
    else:
        upd_total, upd_e2e_total, launches_total = upd_local, upd_e2e, launches
        ms_sweep_max, ms_chem_max, ms_ar_max = ms_sweep, ms_chem, ms_ar

    if rank == 0:
        value = upd_total / (ms * 1e-3)
        peak, peak_src = measured_peaks()
        # dominant kernel: k_sweep_shell.  achieved = algorithmic bytes (104 B x updates of this rank) / its device time.
        sweep_gbs = BYTES_PER_UPDATE * upd_local / (ms_sweep * 1e-3) / 1e9
        fp64 = c.measure_fp64()
        roofline = {"kernel": "k_sweep_shell", "bound": "hbm", "achieved": sweep_gbs, "peak": peak, "unit": "GB/s",
                    "frac": sweep_gbs / peak, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "traffic_note": "ncu capture of one r=56 launch (602 k updates): "
                    f"{NCU_TRAFFIC_BYTES_PER_UPDATE:.0f} B/update measured vs {BYTES_PER_UPDATE} B algorithmic (64-byte DRAM granules on the strided "
                    "x-faces of a shell for the 80-byte cell records and the rate-grid atomics; shell scratch)", "peak_source": peak_src,
                    "launches": int(launches), "avg_launch_ms": ms_sweep / max(1, sum(s["niter"] for s in stats)) ,
                    "note": "the sweep is FP64/LSU bound, not HBM bound (SURVEY F6): see fp64 fields",
                    "updates_per_s_kernel": upd_local / (ms_sweep * 1e-3), "fp64_peak_tflops_measured": fp64,
                    "chem_cells_per_s": sum(s["chem_cells"] for s in stats) / (ms_chem * 1e-3),
                    "chem_hbm_frac": 224 * sum(s["chem_cells"] for s in stats) / (ms_chem * 1e-3) / 1e9 / peak,
                    "ms_sweep": ms_sweep, "ms_chem": ms_chem, "ms_allreduce": ms_ar}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config_dict(p, n_gpus), "clocks": clocks,
                "e2e": {"value": upd_e2e_total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps, "s_per_timestep": t_e2e / e2e_steps},
                "gpu_launches": int(launches_total), "roofline": roofline,
                "niter_per_step": stats[0]["niter"], "s_per_timestep": ms * 1e-3 / args.steps}
        if n_gpus == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            nthreads = O.num_threads()
            ns = min(len(p["NormFlux"]), max(1, nthreads))
            g, ns = cpu_reference_iteration(p, nthreads, sources=ns)
            t0 = time.perf_counter()
            upd, _, _, _ = g.pass_all_sources(nthreads=nthreads, order=0)
            t_rt = time.perf_counter() - t0
            t0 = time.perf_counter()
            g.global_pass(p["dt"], nthreads=nthreads)
            t_ch = time.perf_counter() - t0
            # scale the chemistry share to the full source count so the ratio of sweeps to chemistry matches the workload
            frac = ns / len(p["NormFlux"])
            line["cpu_baseline"] = {"value": upd / (t_rt + t_ch * frac), "unit": UNIT, "cores": nthreads, "kind": "port",
                                    "sample": f"RT pass over {ns} of the {len(p['NormFlux'])} sources ({upd} updates, {t_rt:.1f} s) + "
                                              f"{frac:.2f} x one global chemistry pass ({t_ch:.1f} s) of the same {args.mesh}^3 workload; "
                                              "C++ restatement of the reference (oracle/), OpenMP over sources",
                                    "rt_updates_per_s": upd / t_rt, "chem_cells_per_s": N3 / t_ch}
        print(json.dumps(line))
    c.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mesh", type=int, default=MESH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
