"""BASELINE configs[4]: the chemistry-only global pass (doric + thermal) over 1024^3 cells as 8 slabs of 1024 x 1024 x 128.

  python tools/bench_chem_slabs.py [--slabs 8] [--n 1024] [--iso]          the 8 slabs one after the other on one GPU
  torchrun --nproc-per-node 8 tools/bench_chem_slabs.py [--iso]            one slab per GPU (cells are independent: no
                                                                            data-path collective, weak scaling)
Slab s uses the synthetic inputs of SURVEY 8d with seed 5+s.  Times are CUDA-event times of the pass itself (inputs resident in
HBM); under torchrun the slowest rank's time counts.  One JSON line."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import c2ray_b200


def run_slab(n, seed, iso, device, reps):
    mesh = [n, n, max(1, n // 8)]
    ncell = int(np.prod(mesh))
    q = c2ray_b200.synth.make_chemistry_problem(ncell, seed=seed, isothermal=iso)
    par = c2ray_b200.C2RayParameters(isothermal=iso, H0=q["H0"], Omega0=q["Omega0"])
    c = c2ray_b200.C2Ray(mesh, par, device=device)
    c.setup_cool()
    c.set_geometry([1e22] * 3, 1e66, q["zred"])
    c.set_state(q["ndens"], q["xh"], q["xhe"], q["temperature_grid"])
    c.snapshot_state()
    c.set_rates(q["phih"], q["phihe"], q["phiheat"])
    c.bench_global_pass(q["dt"], 1)          # untimed: decides which kernel the next pass takes
    ms, cf = c.bench_global_pass(q["dt"], reps)
    c.close()
    return ncell, ms, int(cf)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slabs", type=int, default=8)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--iso", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = 200 if a.iso else 224
    if world == 1:
        rows = [run_slab(a.n, 5 + s, a.iso, 0, a.reps) for s in range(a.slabs)]
        cells = sum(r[0] for r in rows); ms = sum(r[1] for r in rows)
        out = {"what": f"{a.slabs} slabs of {a.n}x{a.n}x{a.n // 8} cells, one GPU, one after the other", "ms_per_slab": [r[1] for r in rows]}
    else:
        import torch, torch.distributed as dist
        rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ncell, ms_r, cf = run_slab(a.n, 5 + rank, a.iso, local, a.reps)
        t = torch.tensor([ms_r], dtype=torch.float64, device="cuda")
        tl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl, t)
        ms = max(float(x.item()) for x in tl); cells = ncell * world
        out = {"what": f"one slab of {a.n}x{a.n}x{a.n // 8} cells per GPU on {world} GPUs", "ms_per_rank": [float(x.item()) for x in tl]}
        dist.destroy_process_group()
        if rank != 0:
            return
    out.update({"variant": "isothermal" if a.iso else "thermal", "cells": cells, "ms_total": ms, "chem_cells_per_s": cells / (ms * 1e-3),
                "bytes_per_cell": B, "achieved_gbs": B * cells / (ms * 1e-3) / 1e9, "n_gpus": world})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
