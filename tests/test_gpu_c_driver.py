"""The drop-in boundary from a compiled host: tests/c_driver/evolve3d_driver.c (plain C, the calls the Fortran shim
makes) is compiled against include/c2ray_b200.h, linked with libc2ray_b200.so and run on the same problem as the Python
mirror; in deterministic mode the two must agree bit for bit."""
import os
import struct
import subprocess

import numpy as np
import pytest

import c2ray_b200
from c2ray_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_driver", "evolve3d_driver.c")


def build_driver(out):
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["/usr/bin/gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", out,
                           "-L", libdir, "-l:libc2ray_b200.so", f"-Wl,-rpath,{libdir}"])


def test_c_driver_compiles_and_links(tmp_path):
    """CPU part: the header is valid C and every call the driver makes resolves against the library."""
    build_driver(str(tmp_path / "drv"))
    r = subprocess.run([str(tmp_path / "drv")], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("iso", [False, True])
def test_c_driver_matches_python_mirror(iso, tmp_path):
    p = synth.make_problem(2, n=20, num_src=3, isothermal=iso)
    p["NormFlux"] = p["NormFlux"] * 30.0
    logT, logL = c2ray_b200.read_cooling_tables()
    blob = tmp_path / "problem.bin"
    with open(blob, "wb") as f:
        m = p["mesh"]
        f.write(struct.pack("<8i", int(m[0]), int(m[1]), int(m[2]), len(p["NormFlux"]), int(iso), int(p["cosmological"]),
                            int(p["subboxsize"]), int(p["max_subbox"])))
        f.write(struct.pack("<14d", p["temper_val"], p["H0"], p["Omega0"], p["clumping"], p["T_eff"], p["S_star"],
                            *[float(x) for x in p["dr"]], p["vol"], p["zred"], p["dt"], 0.0, 0.0))
        for a, dt in ((logT, np.float64), (logL, np.float64), (p["srcpos"], np.int32), (p["NormFlux"], np.float64),
                      (p["ndens"], np.float64), (p["xh"], np.float64), (p["xhe"], np.float64), (p["temperature_grid"], np.float32)):
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())
    drv = str(tmp_path / "drv")
    build_driver(drv)
    out = tmp_path / "result.bin"
    r = subprocess.run([drv, str(blob), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    raw = open(out, "rb").read()
    niter, conv_flag, upd, loss = struct.unpack_from("<iiqd", raw, 0)
    n3 = 20 ** 3
    off = 24
    xh = np.frombuffer(raw, np.float64, 2 * n3, off); off += 16 * n3
    xhe = np.frombuffer(raw, np.float64, 3 * n3, off); off += 24 * n3
    T = np.frombuffer(raw, np.float32, 3 * n3, off); off += 12 * n3
    phih = np.frombuffer(raw, np.float64, n3, off)
    # the same step through the Python mirror (device rad_ini, deterministic)
    c = c2ray_b200.from_problem(p, deterministic=True)
    xh0, xhe0, T0 = p["xh"].copy(), p["xhe"].copy(), p["temperature_grid"].copy()
    s = c.evolve3D_host(0.0, p["dt"], 0, p["ndens"], xh0, xhe0, None if iso else T0)
    assert (niter, conv_flag, upd) == (s["niter"], s["conv_flag"], s["rt_updates"])
    assert abs(loss - s["photon_loss_all"]) <= 1e-12 * abs(loss)  # summed with atomics within a shell: order varies
    assert np.array_equal(xh, xh0.ravel()) and np.array_equal(xhe, xhe0.ravel())
    if not iso:
        assert np.array_equal(T, T0.ravel())
    assert np.array_equal(phih, c.get_rates()[0].ravel())
    assert niter >= 2 and xh.reshape(2, -1)[1].max() > 1e-3
    c.close()
