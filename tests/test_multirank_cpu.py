"""world_size-2 (and 3) test of the source-sharded path on CPU with the gloo backend: every rank traces its
do_grid_static share of the sources (master_slave.F90:85) with the oracle, the rate grids + photon-loss + sub-box counters
are summed with one all_reduce over the same contiguous buffer layout the GPU path uses
([phih | phihe(0) | phihe(1) | phiheat | photon_loss(47) | sum_nbox], evolve.F90:505-548), and the result must equal the
single-rank pass.  This exercises the host-side partition/reduction logic; the NCCL path itself runs in bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import c2ray_b200
from oracle import oracle as O
from common import oracle_setup, oracle_grid, relerr

synth = c2ray_b200.synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = synth.make_problem(3, n=12, num_src=5)
    oracle_setup(p)
    g = oracle_grid(p)
    g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"])
    g.set_rates_to_zero()
    upd, nbox, loss, sum_nbox = g.pass_all_sources(rank=rank, npr=world)
    mine = c2ray_b200.source_partition(len(p["NormFlux"]), rank, world)
    assert [i + 1 for i in range(len(nbox)) if nbox[i] > 0] == mine
    phih, phihe, phiheat = g.get_rates()
    N3 = phih.size
    buf = np.zeros(4 * N3 + 48)
    buf[:N3] = phih.ravel(); buf[N3:3 * N3] = phihe.ravel(); buf[3 * N3:4 * N3] = phiheat.ravel()
    buf[4 * N3] = loss; buf[4 * N3 + 47] = sum_nbox
    t = torch.from_numpy(buf)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    u = torch.tensor([upd], dtype=torch.int64)
    dist.all_reduce(u)
    # replicated global pass on the reduced rates (evolve.F90:477-484 runs on every rank)
    g.set_rates(buf[:N3], buf[N3:3 * N3], buf[3 * N3:4 * N3])
    cf = g.global_pass(p["dt"])
    xh_av = g.get_work_state()[0]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), buf=buf, upd=u.numpy(), cf=cf, xh_av=xh_av)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_source_sharded_pass_matches_single_rank(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    p = synth.make_problem(3, n=12, num_src=5)
    oracle_setup(p)
    g = oracle_grid(p)
    g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"])
    g.set_rates_to_zero()
    upd, nbox, loss, sum_nbox = g.pass_all_sources()
    phih, phihe, phiheat = g.get_rates()
    cf = g.global_pass(p["dt"])
    N3 = phih.size
    ranks = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for r in ranks:
        buf = r["buf"]
        assert int(r["upd"][0]) == upd
        assert relerr(buf[:N3], phih.ravel(), 1e-300) < 1e-12            # summation order differs (evolve.F90:523-541)
        assert relerr(buf[N3:3 * N3], phihe.ravel(), 1e-300) < 1e-12
        assert relerr(buf[3 * N3:4 * N3], phiheat.ravel(), 1e-6 * np.abs(phiheat).max()) < 1e-9
        assert abs(buf[4 * N3] / loss - 1) < 1e-12 and int(round(buf[4 * N3 + 47])) == sum_nbox
        assert int(r["cf"]) == cf
        assert np.array_equal(r["buf"], ranks[0]["buf"])                 # every rank holds the same reduced grids
        assert np.array_equal(r["xh_av"], ranks[0]["xh_av"])


def _worker_split(rank, world, port, out_dir):
    """Two global iterations of the scheme c2ray_b200_evolve3d uses with a communicator: every rank traces its sources,
    the rate grids are reduce-scattered (here: all_reduce, then only the own chunk is used), the global pass runs on the
    rank's N3/npr cells, the convergence counters are summed and the fractions all-gathered for the next sweep."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = synth.make_problem(2, n=12, num_src=5)
    p["NormFlux"] = p["NormFlux"] * 30.0
    oracle_setup(p)
    g = oracle_grid(p)
    g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"])
    N3 = 12 ** 3
    chunk = N3 // world
    p0, p1 = rank * chunk, (rank + 1) * chunk
    hist = []
    for it in range(2):
        g.set_rates_to_zero()
        g.pass_all_sources(rank=rank, npr=world)
        phih, phihe, phiheat = g.get_rates()
        buf = torch.from_numpy(np.concatenate([phih.ravel(), phihe.ravel(), phiheat.ravel()]))
        dist.all_reduce(buf)                                   # reduce-scatter + the part of it this rank does not need
        b = buf.numpy()
        mine = np.zeros_like(b)                                # a rank only ever looks at the summed rates of its own cells
        for q in range(4):
            mine[q * N3 + p0:q * N3 + p1] = b[q * N3 + p0:q * N3 + p1]
        g.set_rates(mine[:N3], mine[N3:3 * N3], mine[3 * N3:])
        cf = torch.tensor([g.global_pass_range(p["dt"], p0, p1)], dtype=torch.int64)
        dist.all_reduce(cf)
        hist.append(int(cf))
        work = [w.reshape(w.shape[0], -1) for w in g.get_work_state()]
        gathered = []
        for w in work:                                          # all-gather of every rank's chunk of every plane
            own = torch.from_numpy(np.ascontiguousarray(w[:, p0:p1]))
            parts = [torch.empty_like(own) for _ in range(world)]
            dist.all_gather(parts, own)
            gathered.append(np.concatenate([x.numpy() for x in parts], axis=1))
        T = g.get_state()[2].reshape(3, -1)
        ownT = torch.from_numpy(np.ascontiguousarray(T[:, p0:p1]))
        partsT = [torch.empty_like(ownT) for _ in range(world)]
        dist.all_gather(partsT, ownT)
        Tfull = np.concatenate([x.numpy() for x in partsT], axis=1)
        shp = (12, 12, 12)
        g.set_work_state(*[w.reshape((w.shape[0],) + shp) for w in gathered])
        g.set_state(p["ndens"], p["xh"], p["xhe"], Tfull.reshape((3,) + shp))
    np.savez(os.path.join(out_dir, f"split{rank}.npz"), hist=np.array(hist), xh_av=gathered[0], xhe_av=gathered[1],
             xh_int=gathered[2], xhe_int=gathered[3], T=Tfull)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_split_global_pass_matches_replicated(world, tmp_path):
    mp.spawn(_worker_split, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    p = synth.make_problem(2, n=12, num_src=5)
    p["NormFlux"] = p["NormFlux"] * 30.0
    oracle_setup(p)
    g = oracle_grid(p)
    g.set_work_state(p["xh"], p["xhe"], p["xh"], p["xhe"])
    hist = []
    for it in range(2):
        g.set_rates_to_zero()
        g.pass_all_sources()
        hist.append(g.global_pass(p["dt"]))
    ref = [w.reshape(w.shape[0], -1) for w in g.get_work_state()]
    Tref = g.get_state()[2].reshape(3, -1)
    for r in range(world):
        f = np.load(os.path.join(tmp_path, f"split{r}.npz"))
        assert list(f["hist"]) == hist
        for name, a in zip(("xh_av", "xhe_av", "xh_int", "xhe_int"), ref):
            # the rates are summed in another order (rank by rank): doric turns 1e-16 into ~1e-11, tests/common.py
            assert np.max(np.abs(f[name] - a) / (1e-8 * np.abs(a) + 2e-10)) < 1, (r, name)
        assert relerr(f["T"][:2], Tref[:2]) < 1.3e-7
        g0 = np.load(os.path.join(tmp_path, "split0.npz"))
        assert all(np.array_equal(f[k], g0[k]) for k in ("xh_av", "xhe_av", "xh_int", "xhe_int", "T"))   # every rank: same state
