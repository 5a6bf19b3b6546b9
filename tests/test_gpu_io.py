"""GPU tests of the iteration dump / restart path (evolve.F90:233-367) and of output streams 2 and 3
(output.F90:249-379): the files are read back with scipy's independent Fortran-record reader, and a time step resumed
from a dump must end exactly where the uninterrupted step ended."""
import os

import numpy as np
import pytest
from scipy.io import FortranFile

import c2ray_b200
from c2ray_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _ctx(p, monkeypatch):
    monkeypatch.setenv("C2RAY_CHEM_QUEUE", "0")  # one global-pass kernel throughout: bitwise comparisons below
    return c2ray_b200.from_problem(p, deterministic=True)


@pytest.mark.parametrize("iso", [False, True])
def test_dump_contents_and_restart(iso, tmp_path, monkeypatch):
    p = synth.make_problem(2, n=20, num_src=3, isothermal=iso)
    n3 = 20 ** 3
    # uninterrupted step
    c = _ctx(p, monkeypatch)
    s_full = c.evolve3D(0.0, p["dt"], 0)
    ref = c.get_state() + tuple(c.get_rates())
    c.close()
    assert s_full["niter"] >= 3

    # the same step with a dump after every source pass (interval 0 s): the files alternate iterdump1/iterdump2
    d = str(tmp_path) + "/"
    c = _ctx(p, monkeypatch)
    c.set_dump(d, 0.0)
    s_dump = c.evolve3D(0.0, p["dt"], 0)
    assert s_dump["niter"] == s_full["niter"]
    for a, b in zip(c.get_state() + tuple(c.get_rates()), ref):
        assert np.array_equal(a, b)
    c.close()
    last = s_full["niter"]
    name = {1: "iterdump1.bin", 0: "iterdump2.bin"}
    files = {k: d + name[k % 2] for k in (last - 1, last)}
    f = FortranFile(files[last], "r")
    assert f.read_ints(np.int32)[0] == last
    loss = f.read_reals(np.float64)
    assert loss.shape == (47,) and loss[0] == pytest.approx(s_full["photon_loss_all"], rel=1e-12) and not loss[1:].any()
    sizes = [n3, 2 * n3, 2 * n3, 2 * n3, 3 * n3, 3 * n3] + ([] if iso else [n3])
    recs = [f.read_reals(np.float64) for _ in sizes]
    assert [r.size for r in recs] == sizes
    if not iso:
        assert f.read_reals(np.float32).size == 3 * n3
    with pytest.raises(Exception):
        f.read_ints(np.int32)  # no further record
    f.close()
    assert np.array_equal(recs[0].reshape(20, 20, 20), ref[3])           # phih_grid of the last iteration
    assert np.array_equal(recs[3].reshape(2, 20, 20, 20), ref[4])        # phihe_grid

    # resume from the dump of the last-but-one iteration: start_from_dump + global_pass, then the loop goes on
    which = 1 if (last - 1) % 2 == 1 else 2
    f = FortranFile(files[last - 1], "r")
    assert f.read_ints(np.int32)[0] == last - 1
    f.close()
    c = _ctx(p, monkeypatch)
    c.set_dump(d, -1.0)
    s_res = c.evolve3D(0.0, p["dt"], which)
    assert s_res["niter"] == s_full["niter"] and s_res["conv_flag"] == s_full["conv_flag"]
    for a, b in zip(c.get_state() + tuple(c.get_rates()), ref):
        assert np.array_equal(a, b)
    # explicit-path variants, and restart=3 -> iterdump.bin
    c.write_iteration_dump(d + "iterdump.bin", 5)
    assert c.start_from_dump(d + "iterdump.bin") == 5
    with pytest.raises(capi.C2RayError, match="cannot open"):
        c.start_from_dump(d + "nothing.bin")
    c.close()
    # a dump of another mesh is rejected, not misread
    q = synth.make_problem(2, n=16, num_src=3, isothermal=iso)
    c = _ctx(q, monkeypatch)
    with pytest.raises(capi.C2RayError, match="does not match"):
        c.start_from_dump(files[last])
    c.close()


def test_output_streams(tmp_path, monkeypatch):
    p = synth.make_problem(2, n=16, num_src=2, isothermal=False)
    c = _ctx(p, monkeypatch)
    c.evolve3D(0.0, p["dt"], 0)
    xh, xhe, T = c.get_state()
    phih, phihe, phiheat = c.get_rates()
    d = str(tmp_path)
    c.write_stream2(d, 8.85)
    c.write_stream3(d, 8.85)
    c.close()

    def read(name, dtype):
        f = FortranFile(os.path.join(d, name), "r")
        m = f.read_ints(np.int32)
        a = f.read_reals(dtype)
        f.close()
        assert list(m) == [16, 16, 16]
        return a.reshape(16, 16, 16)

    assert np.array_equal(read("xfrac3d_8.850.bin", np.float64), xh[1])
    assert np.array_equal(read("xfrac3dHe1_8.850.bin", np.float64), xhe[1])
    assert np.array_equal(read("xfrac3dHe2_8.850.bin", np.float64), xhe[2])
    assert np.array_equal(read("Temper3D_8.850.bin", np.float32), T[0])
    assert np.array_equal(read("IonRates3D_8.850.bin", np.float32), phih.astype(np.float32))
    assert np.array_equal(read("HeatRates3D_8.850.bin", np.float32), phiheat.astype(np.float32))
    assert phih.max() > 0 and phiheat.max() > 0
