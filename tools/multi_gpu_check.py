"""Multi-GPU parity check, run under torchrun with N >= 2 ranks (one per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/multi_gpu_check.py [mesh] [num_src] [steps]

For each of two consecutive evolve3D time steps it compares
  (a) the split global pass (reduce-scatter -> chemistry on N3/npr cells -> all-gather, the default with a communicator)
      against the reference's scheme (allreduce of the rate grids + replicated pass, C2RAY_SPLIT_CHEM=0): bitwise at 2
      ranks in deterministic mode (a+b is the only sum), 1e-12 otherwise;
  (b) both against one rank tracing every source alone (tolerance: summation order of the rate grids);
  (c) that every rank ends with the same full state and rate grids;
  (d) the balanced source schedule against the static one, and a dump written under the split pass resuming bitwise;
  (e) the NCCL-reduced rate grids of one source pass against the CPU oracle: every rank's own update count equals
      orc_pass_all_sources(rank, npr) (do_grid_static, master_slave.F90:85), the reduced grids equal the oracle's pass
      over all sources to 1e-8 relative (evolve.F90:505-548).
Prints one line per rank and exits non-zero on a mismatch.  tests/test_gpu_multirank.py runs it under pytest."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import c2ray_b200


def run(p, local, uid, rank, world, split, steps, schedule=0):
    os.environ["C2RAY_SPLIT_CHEM"] = "1" if split else "0"
    c = c2ray_b200.from_problem(p, device=local, deterministic=True)
    if uid is not None:
        c.comm_init(uid, rank, world)
        c.set_source_schedule(schedule)
    hist = []
    for s in range(steps):
        st = c.evolve3D(0.0, p["dt"], 0)
        hist.append((st["niter"], tuple(int(x) for x in st["conv_hist"]), st["sum_nbox_all"], st["photon_loss_all"],
                     st["photcons"], st["ms_chem"], st["ms_allreduce"], st["ms_sweep"]))
    out = c.get_state() + tuple(c.get_rates()) + tuple(c.get_work_state())
    mine = list(c.my_sources())
    c.close()
    return hist, out + (mine,)


def run_dump_restart(p, local, uid_fn, rank, world, tmpdir):
    """Split pass + iteration dumps after every source pass (rank 0 writes), then a fresh set of contexts resumes from
    the last-but-one dump: must end bitwise where the uninterrupted run ended (deterministic sweeps, one chemistry kernel)."""
    os.environ["C2RAY_SPLIT_CHEM"] = "1"
    os.environ["C2RAY_CHEM_QUEUE"] = "0"
    c = c2ray_b200.from_problem(p, device=local, deterministic=True)
    c.comm_init(uid_fn(), rank, world)
    c.set_dump(tmpdir + "/", 0.0)
    st = c.evolve3D(0.0, p["dt"], 0)
    ref = c.get_state() + tuple(c.get_rates())
    c.close()
    dist.barrier()
    last = st["niter"]
    which = 1 if (last - 1) % 2 == 1 else 2
    c = c2ray_b200.from_problem(p, device=local, deterministic=True)
    c.comm_init(uid_fn(), rank, world)
    c.set_dump(tmpdir + "/", -1.0)
    st2 = c.evolve3D(0.0, p["dt"], which)
    out = c.get_state() + tuple(c.get_rates())
    c.close()
    os.environ.pop("C2RAY_CHEM_QUEUE")
    same = st2["niter"] == last and all(np.array_equal(a, b) for a, b in zip(out, ref))
    return same, last


def relerr(a, b):
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-300)))


def oracle_check(p, local, uid, rank, world):
    """(e): one source pass from a partially ionized state; NCCL allreduce vs the oracle."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from common import oracle_setup, oracle_grid, partially_ionized_state
    tables = oracle_setup(p)
    xh, xhe = partially_ionized_state(p, seed=3)
    c = c2ray_b200.from_problem(p, device=local, tables=tables)
    c.comm_init(uid, rank, world)
    c.begin_step()
    c.set_work_state(xh, xhe, xh, xhe)
    c.set_rates_to_zero()
    upd_mine = c.pass_all_sources(1, p["dt"])
    rates = c.get_rates()
    c.close()
    g = oracle_grid(p)
    g.set_work_state(xh, xhe, xh, xhe)
    g.set_rates_to_zero()
    upd_o_mine = g.pass_all_sources(rank=rank, npr=world)[0]
    g.set_rates_to_zero()
    upd_o_all = g.pass_all_sources()[0]
    ref = g.get_rates()
    err = 0.0
    same_zero = True
    for a, b in zip(rates, ref):
        nz = b != 0.0
        same_zero &= bool(np.array_equal(a != 0.0, nz))
        if nz.any():
            err = max(err, float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz]))))
    tot = torch.tensor([upd_mine], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    ok = upd_mine == upd_o_mine and int(tot.item()) == upd_o_all and same_zero and err < 1e-8
    return ok, err, upd_mine, upd_o_mine


def main():
    mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = c2ray_b200.synth.make_problem(2, n=mesh, num_src=nsrc, isothermal=False)
    p["subboxsize"] = 5  # sub-box counts (the balanced schedule's cost) then differ from source to source

    def uid():
        u = [c2ray_b200.C2Ray.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(u, src=0)
        return u[0]

    h_split, s_split = run(p, local, uid(), rank, world, True, steps)
    h_repl, s_repl = run(p, local, uid(), rank, world, False, steps)
    h_one, s_one = run(p, local, None, 0, 1, False, steps)
    h_bal, s_bal = run(p, local, uid(), rank, world, True, steps, schedule=1)
    mine_static, mine_bal = s_split[-1], s_bal[-1]
    s_split, s_repl, s_one, s_bal = s_split[:-1], s_repl[:-1], s_one[:-1], s_bal[:-1]
    import tempfile
    tmp = [tempfile.mkdtemp(prefix="c2ray_dump_") if rank == 0 else None]
    dist.broadcast_object_list(tmp, src=0)
    restart_ok, restart_niter = run_dump_restart(p, local, uid, rank, world, tmp[0])
    ok = restart_ok
    oracle_ok, oracle_err, upd_mine, upd_o_mine = oracle_check(p, local, uid(), rank, world)
    ok &= oracle_ok
    # the balanced schedule: same integer histories, same fields to summation order, every source dealt exactly once
    counts = torch.zeros(len(p["NormFlux"]) + 1, dtype=torch.int32, device="cuda")
    counts[torch.tensor(mine_bal, dtype=torch.long, device="cuda")] += 1
    dist.all_reduce(counts)
    ok &= bool((counts[1:] == 1).all().item()) and int(counts[0].item()) == 0
    for a, b in zip(h_bal, h_one):
        ok &= a[:3] == b[:3]
    e_bal = max(float(np.max(np.abs(x - y) / (1e-8 * np.abs(y) + 2e-10))) for x, y in zip(s_bal[:2], s_one[:2]))
    ok &= e_bal < 1
    # integer histories: niter, conv_flag per iteration, sum_nbox
    for a, b, c1 in zip(h_split, h_repl, h_one):
        ok &= a[:3] == b[:3] == c1[:3]
        ok &= abs(a[3] - c1[3]) <= 1e-10 * abs(c1[3]) + 1e-300 and abs(a[4] - c1[4]) <= 1e-9
    # two ranks: a+b is the only sum, reduce-scatter and allreduce must agree bitwise.  More ranks: NCCL may add in a
    # different order (1e-16 on the rate grids), which doric's cancellation noise turns into ~1e-11 absolute on the
    # fractions (tests/common.py) -- compared with the parity tolerances, in units of them
    def tol_err(n, x, y):
        if n in ("xh", "xhe", "xh_av", "xhe_av", "xh_int", "xhe_int"):
            return float(np.max(np.abs(x - y) / (1e-8 * np.abs(y) + 2e-10)))
        if n == "T":
            return float(np.max(np.abs(x.astype(np.float64) - y) / (1.3e-7 * np.abs(y) + 1e-30)))
        return float(np.max(np.abs(x - y) / (1e-8 * np.abs(y) + 1e-300)))
    names = ["xh", "xhe", "T", "phih", "phihe", "phiheat", "xh_av", "xhe_av", "xh_int", "xhe_int"]
    if world == 2:
        tol_sr = 0.0
        e_sr = max(relerr(x, y) for x, y in zip(s_split, s_repl))
        ok &= e_sr <= tol_sr
    else:
        tol_sr = 1.0
        e_sr = max(tol_err(n, x, y) for n, x, y in zip(names, s_split, s_repl))
        ok &= e_sr < tol_sr
    # against one rank: fractions 1e-8 rel + 2e-10 abs (tests/common.py), T one float ulp, rates 1e-8
    errs = {n: tol_err(n, x, y) for n, x, y in zip(names, s_split, s_one)}
    ok &= all(v < 1 for v in errs.values())
    # every rank holds the same full result
    mine = torch.tensor([float(np.sum(a.astype(np.float64))) for a in s_split], dtype=torch.float64, device="cuda")
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    ok &= same
    ms = lambda h: (sum(x[7] for x in h), sum(x[5] for x in h), sum(x[6] for x in h))
    print(f"rank {rank}/{world}: mesh {mesh} sources {nsrc} niter {[h[0] for h in h_split]} split-vs-replicated {e_sr:.2e} "
          f"(tol {tol_sr:g}) vs-one-rank err/tol {max(errs.values()):.3f} all-ranks-equal {same} | ms sweep/chem/comm split "
          f"{ms(h_split)[0]:.1f}/{ms(h_split)[1]:.1f}/{ms(h_split)[2]:.1f} replicated {ms(h_repl)[0]:.1f}/{ms(h_repl)[1]:.1f}/{ms(h_repl)[2]:.1f}"
          f" | balanced schedule: sources {[int(x) for x in mine_static]} -> {[int(x) for x in mine_bal]}, err/tol {e_bal:.3f}"
          f" | dump+restart under the split pass (niter {restart_niter}) bitwise {restart_ok}"
          f" | NCCL-reduced rates vs oracle: max rel {oracle_err:.2e}, own updates {upd_mine} == oracle(rank,npr) {upd_o_mine}: {oracle_ok}"
          f" -> {'OK' if ok else 'MISMATCH'}", flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
