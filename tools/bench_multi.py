"""A BASELINE config on N GPUs of one node (torchrun, one rank per GPU): full evolve3D time steps, sources sharded over
the ranks, split global pass.  BASELINE configs[3] is `bench_multi.py 4 512 10000`:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
      tools/bench_multi.py <config 1..4> <mesh> <num_src> [steps] [schedule 0|1]

Every rank builds the same synthetic problem, rank r traces its share of the sources; device times are CUDA-event times
per rank, reduced with MAX over the ranks; one JSON line from rank 0.  The first step is untimed (allocations, NCCL
set-up); consecutive steps continue from the evolving state, as the program's time loop does."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import c2ray_b200


def main():
    cfg, mesh, nsrc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    schedule = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.time()
    p = c2ray_b200.synth.make_problem(cfg, n=mesh, num_src=nsrc)
    c = c2ray_b200.from_problem(p, device=local)
    if world > 1:
        uid = [c2ray_b200.C2Ray.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        c.comm_init(uid[0], rank, world)
        c.set_source_schedule(schedule)
    setup_s = time.time() - t0
    rows = []
    for step in range(steps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        c.timer_start()
        s = c.evolve3D(step * p["dt"], p["dt"], 0)
        ms = c.timer_stop()
        rows.append((ms, s))
    free, total = torch.cuda.mem_get_info()
    timed = rows[1:]
    loc = np.array([[ms, s["ms_sweep"], s["ms_chem"], s["ms_allreduce"], s["rt_updates"]] for ms, s in timed])
    t = torch.tensor(loc, dtype=torch.float64, device="cuda")
    tmax, tsum = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    # every rank must hold the same final state
    xh = c.get_state()[0]
    chk = torch.tensor([float(xh[1].sum()), float(xh[1].max())], dtype=torch.float64, device="cuda")
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        tm, ts = tmax.cpu().numpy(), tsum.cpu().numpy()
        out = {"config": cfg, "mesh": mesh, "sources": nsrc, "n_gpus": world, "schedule": "balanced" if schedule else "static",
               "setup_s": setup_s, "untimed_first_step_s": rows[0][0] * 1e-3, "device_mem_used_gb": (total - free) / 1e9,
               "ranks_hold_same_state": bool(torch.equal(lo, hi)), "mean_xHII": float(xh[1].mean()),
               "steps": [{"s_per_timestep": tm[i, 0] * 1e-3, "niter": timed[i][1]["niter"], "conv_flag": timed[i][1]["conv_flag"],
                          "rt_updates_all_ranks": int(ts[i, 4]), "updates_per_s": ts[i, 4] / (tm[i, 0] * 1e-3),
                          "ms_sweep_max": tm[i, 1], "ms_sweep_mean": ts[i, 1] / world, "ms_chem_max": tm[i, 2],
                          "ms_collectives_max": tm[i, 3], "photcons": timed[i][1]["photcons"]} for i in range(len(timed))]}
        print(json.dumps(out))
    c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
