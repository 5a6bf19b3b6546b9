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
