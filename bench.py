#!/usr/bin/env python
"""bench.py -- headline benchmark of the C2-Ray H+He hot path on B200 (see DESIGN.md "Measurement").

Metric: source x cell RT updates/s (BASELINE.json), whole job, over full evolve3D time steps (ray-tracing sweeps of all
sources + rate-grid reduction + global chemistry passes, iterated to convergence).
Workload: BASELINE configs[2] -- 256^3 lognormal box, 1000 sources (BB 5e4 K + QPL on the 50 brightest), subboxsize 10,
non-isothermal -- the largest source-sharded configuration that fits one GPU.  STRONG scaling: for N > 1 the same 1000
sources are dealt over the ranks (balanced schedule); per iteration the rate grids are reduce-scattered, every rank runs
the global pass on its N^3/npr cells and the fractions the next sweep reads are all-gathered.
A step = one evolve3D(time,dt) from the same start state (device snapshot restored inside the timed region).
`--workload config1` selects round 1's headline (BASELINE configs[1], 128^3, 16 sources per GPU, weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload config2|config1]
  torchrun ... bench.py --gpus N ...        (one rank per GPU)

--impl reference times the CPU restatement of the reference (oracle/, all host threads of rank 0) on a bounded sample
of the same workload, see CpuSample.
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
BYTES_PER_UPDATE = 104  # 5 FP64 state reads + read-modify-write of 4 rate grids (thermal)

# Workloads.  "config2" (default) is BASELINE configs[2] -- the largest configuration that is source-sharded AND fits one
# GPU: 256^3 lognormal box, 1000 sources (black body 5e4 K; the 50 brightest also emit a hard quasar-like power law),
# subboxsize 10, non-isothermal.  It is STRONG-scaled: the same 1000 sources are dealt over the N GPUs.
# "config1" is BASELINE configs[1] (128^3 Test-4 style, 16 BB sources per GPU, weak scaling): round 1's headline, kept
# for comparison (`--workload config1`).
WORKLOADS = {
    "config2": dict(config=3, mesh=256, sources=1000, scaling="strong", cpu_sample_sources=2,
                    label="BASELINE configs[2]: 256^3 lognormal box (sigma=1.2, seed 256), 1000 sources at density peaks "
                          "(BB T_eff=5e4 K; QPL index 1.8 on the 50 brightest), subboxsize=10, non-isothermal, dt=5 Myr"),
    "config1": dict(config=2, mesh=128, sources=16, scaling="weak", cpu_sample_sources=16,
                    label="BASELINE configs[1]: Test-4-style 128^3 lognormal box (sigma=1, seed 4), 16 BB sources per GPU "
                          "T_eff=1e5 K, subboxsize=mesh, non-isothermal, dt=0.05 Myr"),
}
# Per-update figures of the dominant kernel taken from the committed ncu captures (profiles/, `tools/extract_ncu.py`):
# FP64 flops (2 DFMA + DADD + DMUL thread instructions) and DRAM bytes per source x cell update.
NCU_FIGURES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "sweep_kernel_figures.json")


def workload(name, n_gpus, mesh=None):
    import c2ray_b200
    w = WORKLOADS[name]
    n = mesh or w["mesh"]
    if w["scaling"] == "strong":
        return c2ray_b200.synth.make_problem(w["config"], n=n, num_src=w["sources"], isothermal=False)
    p = c2ray_b200.synth.make_problem(w["config"], n=n, num_src=w["sources"] * n_gpus, isothermal=False)
    if n_gpus > 1:  # keep every source as bright as in the 16-source box
        p["NormFlux"] = p["NormFlux"] * n_gpus
    return p


def config_dict(name, p, n_gpus):
    w = WORKLOADS[name]
    ns = int(len(p["NormFlux"]))
    if n_gpus == 1:
        par = "one GPU"
    elif w["scaling"] == "strong":
        par = (f"the {ns} sources dealt over {n_gpus} GPUs (balanced by the previous pass's cost, master_slave.F90 analogue); per "
               "iteration one reduce-scatter of the rate grids, the global pass on N^3/npr cells per rank, one all-gather of the "
               "fractions the next sweep reads")
    else:
        par = (f"sources round-robin over {n_gpus} GPUs (do_grid_static); per iteration one reduce-scatter of the rate grids, the "
               "global pass on N^3/npr cells per rank, one all-gather of the fractions the next sweep reads")
    return {"workload": w["label"] + "; one full evolve3D time step per step", "name": name, "mesh": int(p["mesh"][0]),
            "sources": ns, "scaling": w["scaling"], "parallelism": par,
            "l2_policy": "per-step working set (state + rate grids + cell records + snapshot, several hundred MB to GB) exceeds "
                         "the 126 MB L2"}


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
class CpuSample:
    """The bounded CPU sample of a workload, on the oracle (the C++ restatement of the reference; the Fortran itself
    cannot be built in this image).  The sample sources are the workload's `k` brightest (for configs[2] these are BB+QPL
    sources that trace the whole box -- the 50 such sources are 99 % of the GPU arm's updates).  Set-up, untimed: global
    iteration 1 of the time step for those k sources from the neutral start state (RT pass + global chemistry pass), so
    that the timed passes see ionized bubbles around the sources like 11 of the GPU arm's 12 iterations do.  One timed
    sample: the RT pass of iteration 2 over the k sources with every host thread busy (sources over threads, the threads
    that leaves idle inside a source's shells), plus the k/NumSrc share of one global chemistry pass."""

    def __init__(self, name, p, nthreads):
        from oracle import oracle as O
        self.O, self.p, self.nthreads = O, p, nthreads
        w = WORKLOADS[name]
        self.ns_all = len(p["NormFlux"])
        self.k = k = min(w["cpu_sample_sources"], self.ns_all)
        O.rad_ini(p["T_eff"], p["S_star"], qpl=p.get("qpl"), isothermal=p["isothermal"])
        O.set_params(p["isothermal"], p["temper_val"], p["clumping"], p["zred"], p["H0"], p["Omega0"], p["cosmological"],
                     p["subboxsize"], p["max_subbox"])
        g = self.g = O.Grid(p["mesh"], p["dr"], p["vol"])
        g.set_state(p["ndens"], p["xh"], p["xhe"], p["temperature_grid"])
        q = p.get("NormFluxQPL")
        g.set_sources(p["srcpos"][:k], p["NormFlux"][:k], None, None if q is None else q[:k])
        g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"])
        g.set_rates_to_zero()
        t0 = time.perf_counter()
        g.pass_all_sources(nthreads=nthreads, order=2)
        t1 = time.perf_counter()
        g.global_pass(p["dt"], nthreads=nthreads)
        self.t_chem = time.perf_counter() - t1          # one global chemistry pass over the whole mesh
        self.t_setup = time.perf_counter() - t0

    def run(self):
        """-> (updates, seconds) of one sample."""
        self.g.set_rates_to_zero()
        t0 = time.perf_counter()
        upd, _, _, _ = self.g.pass_all_sources(nthreads=self.nthreads, order=2)
        t_rt = time.perf_counter() - t0
        return upd, t_rt + self.t_chem * self.k / self.ns_all, t_rt

    def describe(self):
        n = int(self.p["mesh"][0])
        return (f"RT pass of global iteration 2 over the {self.k} brightest of the {self.ns_all} sources + {self.k}/{self.ns_all} of one "
                f"global chemistry pass ({self.t_chem:.2f} s for the {n}^3 mesh), after an untimed iteration 1 of the same sources "
                f"({self.t_setup:.1f} s); C++ restatement of the reference (oracle/, g++ -O2 -fopenmp), {self.nthreads} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    nthreads = O.set_num_threads()   # all CPUs of this process (torchrun exports OMP_NUM_THREADS=1 to its children)
    n_gpus = int(os.environ.get("WORLD_SIZE", args.gpus))
    p = workload(args.workload, n_gpus, args.mesh)   # the same config as the GPU arm at this N
    cs = CpuSample(args.workload, p, nthreads)
    times, updates = [], 0
    for step in range(args.warmup + args.steps):
        upd, dt, _ = cs.run()
        if step >= args.warmup:
            times.append(dt); updates += upd
    total = sum(times)
    value = updates / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": WORKLOADS[args.workload]["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, p, n_gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": "per step: " + cs.describe()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


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

    w = WORKLOADS[args.workload]
    p = workload(args.workload, n_gpus, args.mesh)
    c = c2ray_b200.from_problem(p, device=local)
    if world > 1:
        uid = [c2ray_b200.C2Ray.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        c.comm_init(uid[0], rank, world)
        if w["scaling"] == "strong":
            c.set_source_schedule(1)   # balanced deal from the previous pass's cost records
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
    l0, sl0 = c.launch_count(), c.sweep_launch_count()
    c.timer_start()
    for step in range(args.steps):
        c.restore_state()
        stats.append(c.evolve3D(0.0, p["dt"], 0))
    ms = c.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = c.launch_count() - l0
    sweep_launches = c.sweep_launch_count() - sl0
    upd_local = sum(s["rt_updates"] for s in stats)
    ms_sweep = sum(s["ms_sweep"] for s in stats)
    ms_chem = sum(s["ms_chem"] for s in stats)
    ms_ar = sum(s["ms_allreduce"] for s in stats)
    niter_total = sum(s["niter"] for s in stats)

    # ---- end-to-end arm: host buffers through the Fortran-facing entry point ------------------------------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_ndens, h_xh0, h_xhe0, h_T0 = pin(p["ndens"]), pin(p["xh"]), pin(p["xhe"]), pin(p["temperature_grid"])
    h_xh, h_xhe, h_T = torch.empty_like(h_xh0).pin_memory(), torch.empty_like(h_xhe0).pin_memory(), torch.empty_like(h_T0).pin_memory()
    e2e_steps = max(1, min(args.steps, 2))
    upd_e2e, t_e2e = 0, 0.0
    for step in range(e2e_steps):
        # the call updates xh, xhe, temperature_grid in place: put the start state back first (host-to-host copies of the
        # bench's own scaffolding, outside the timed region), then time the call itself -- H2D of the inputs from pinned
        # host memory, the whole time step, D2H of the results
        h_xh.copy_(h_xh0); h_xhe.copy_(h_xhe0); h_T.copy_(h_T0)
        barrier()
        t0 = time.perf_counter()
        s = c.evolve3D_host(0.0, p["dt"], 0, h_ndens.numpy(), h_xh.numpy(), h_xhe.numpy(), h_T.numpy())
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
        upd_e2e += s["rt_updates"]
    h2d = N3 * (8 + 16 + 24 + 12)
    d2h = N3 * (16 + 24 + 12)

    # ---- reduce over ranks: max time, summed units --------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms, t_e2e, ms_sweep, ms_chem, ms_ar], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e, ms_sweep_max, ms_chem_max, ms_ar_max = t.tolist()
        u = torch.tensor([upd_local, upd_e2e, launches, ms_sweep], dtype=torch.float64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        upd_total, upd_e2e_total, launches_total, ms_sweep_sum = u.tolist()
    else:
        upd_total, upd_e2e_total, launches_total = upd_local, upd_e2e, launches
        ms_sweep_max, ms_chem_max, ms_ar_max, ms_sweep_sum = ms_sweep, ms_chem, ms_ar, ms_sweep

    if rank == 0:
        value = upd_total / (ms * 1e-3)
        peak, peak_src = measured_peaks()
        fig = {}
        if os.path.exists(NCU_FIGURES):
            fig = json.load(open(NCU_FIGURES)).get(args.workload, {})
        # dominant kernel: the ray-tracing sweep.  Its rate on this rank = this rank's updates / its event-timed duration.
        upd_per_s_kernel = upd_local / (ms_sweep * 1e-3)
        sweep_gbs = BYTES_PER_UPDATE * upd_per_s_kernel / 1e9
        fp64_peak = c.measure_fp64()
        flops_per_update = fig.get("fp64_flops_per_update")
        fp64_tflops = flops_per_update * upd_per_s_kernel / 1e12 if flops_per_update else None
        traffic_per_update = fig.get("dram_bytes_per_update")
        launch_updates = upd_local / max(1, sweep_launches)
        # The kernel is bound by the FP64 pipe, not by HBM (~6 kFLOP of FP64 per 104 algorithmic bytes), so the
        # roofline that binds is FP64: achieved = FP64 flops per update (ncu instruction counts of the committed capture) x
        # this run's updates/s of the kernel; peak = the DFMA rate measured on this GPU in this run.  The HBM view the
        # contract asks for (104 B x updates/s over the measured copy bandwidth) is carried next to it.
        roofline = {"kernel": fig.get("kernel", "k_sweep_shell"), "bound": "fp64",
                    "achieved": fp64_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (fp64_tflops / fp64_peak) if fp64_tflops else None,
                    "fp64_achieved_tflops": fp64_tflops, "fp64_peak_tflops_measured": fp64_peak,
                    "fp64_frac": (fp64_tflops / fp64_peak) if fp64_tflops else None,
                    "fp64_flops_per_update": flops_per_update, "figures_source": fig.get("source"),
                    # the unit nearest its own peak in the committed capture (not a throughput roofline: DESIGN section 3)
                    "busiest_unit_in_capture": fig.get("busiest_unit"),
                    "hbm_achieved_gbs": sweep_gbs, "hbm_peak_gbs": peak, "hbm_frac": sweep_gbs / peak, "hbm_peak_source": peak_src,
                    "hbm_algorithmic_bytes_per_update": BYTES_PER_UPDATE,
                    "traffic": (traffic_per_update * launch_updates) if traffic_per_update else None,
                    "traffic_bytes_per_update": traffic_per_update,
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per update of the committed `ncu --set full` capture "
                                    "x this run's mean updates per launch (not re-measured in this run)",
                    "sweep_launches": int(sweep_launches), "avg_launch_ms": ms_sweep / max(1, sweep_launches),
                    "avg_launch_updates": launch_updates, "ms_per_rt_pass": ms_sweep / max(1, niter_total),
                    "updates_per_s_kernel": upd_per_s_kernel,
                    "chem_cells_per_s": sum(s["chem_cells"] for s in stats) / (ms_chem * 1e-3),
                    "chem_hbm_frac": 224 * sum(s["chem_cells"] for s in stats) / (ms_chem * 1e-3) / 1e9 / peak,
                    "ms_sweep": ms_sweep, "ms_chem": ms_chem, "ms_allreduce": ms_ar,
                    "ms_sweep_max_over_ranks": ms_sweep_max, "ms_sweep_mean_over_ranks": ms_sweep_sum / world,
                    "ms_collectives_max_over_ranks": ms_ar_max}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config_dict(args.workload, p, n_gpus), "clocks": clocks,
                "e2e": {"value": upd_e2e_total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps, "s_per_timestep": t_e2e / e2e_steps},
                "gpu_launches": int(launches_total), "roofline": roofline,
                "niter_per_step": stats[0]["niter"], "s_per_timestep": ms * 1e-3 / args.steps}
        if n_gpus == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            nthreads = O.set_num_threads()
            cs = CpuSample(args.workload, p, nthreads)
            upd, dt, t_rt = cs.run()
            line["cpu_baseline"] = {"value": upd / dt, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": cs.describe(),
                                    "rt_updates_per_s": upd / t_rt, "chem_cells_per_s": N3 / cs.t_chem}
        emit(line)
    c.close()
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's original stdout; everything else that lands on fd 1 meanwhile (NCCL prints its
    version banner there from C) has been sent to stderr, so that the driver reads exactly one line."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--mesh", type=int, default=None, help="override the workload's mesh (smoke runs only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
