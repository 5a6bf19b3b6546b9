"""ctypes loader for the CPU parity oracle (oracle/c2ray_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  The product package never imports this module.  Parity unpinned (see the .cpp header).
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, os.environ.get("C2RAY_ORACLE_LIB", "libc2ray_oracle.so"))

NUMTAU, NUMFREQBND, NUMHEATBIN = 2000, 47, 113
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(HERE, "c2ray_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src) or \
            not os.path.exists(os.path.join(HERE, "libc2ray_oracle_fma.so")) or \
            not os.path.exists(os.path.join(HERE, "libc2ray_oracle_assoc.so")):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        L = _lib
        L.orc_coolin.restype = C.c_double
        L.orc_coolin.argtypes = [C.c_double, C.c_double, _dp, _dp, C.c_double]
        L.orc_photon_loss.restype = C.c_double
        L.orc_sum_nbox.restype = C.c_long
        L.orc_pass_all_sources.restype = C.c_long
        L.orc_last_nit_total.restype = C.c_long
        load_cooling()
    return _lib


def read_cooling_table():
    """data/cooling_h_he.tab -> (logT[801], 5 x log10 Lambda[801])"""
    a = np.loadtxt(os.path.join(ROOT, "data", "cooling_h_he.tab"))
    assert a.shape == (801, 6)
    return [np.ascontiguousarray(a[:, i]) for i in range(6)]


def load_cooling():
    t = read_cooling_table()
    _lib.orc_set_cooling(*[x.ctypes.data_as(C.c_void_p) for x in t])


def rad_ini(T_eff=5.0e4, S_star=1e48, pl=None, qpl=None, isothermal=False):
    """pl / qpl: dict(index, minfreq, maxfreq, S_star) or None."""
    L = lib()
    z = dict(index=1.0, minfreq=1.0, maxfreq=2.0, S_star=0.0)
    p = pl or z
    q = qpl or z
    L.orc_rad_ini(C.c_double(T_eff), C.c_double(S_star), C.c_double(p["index"]), C.c_double(p["minfreq"]),
                  C.c_double(p["maxfreq"]), C.c_double(p["S_star"]), C.c_double(q["index"]), C.c_double(q["minfreq"]),
                  C.c_double(q["maxfreq"]), C.c_double(q["S_star"]), C.c_int(int(isothermal)))


def sed_info():
    out = np.zeros(10)
    lib().orc_get_sed_info(out.ctypes.data_as(C.c_void_p))
    return dict(R_star2=out[0], h_over_kT=out[1], pl_scaling=out[2], qpl_scaling=out[3],
                bb=(int(out[4]), int(out[5])), pl=(int(out[6]), int(out[7])), qpl=(int(out[8]), int(out[9])))


def band_data():
    a = [np.zeros(NUMFREQBND) for _ in range(5)]
    lib().orc_get_band_data(*[x.ctypes.data_as(C.c_void_p) for x in a])
    return dict(freq_min=a[0], freq_max=a[1], sigma_HI=a[2], sigma_HeI=a[3], sigma_HeII=a[4])


def romw():
    out = np.zeros(513)
    lib().orc_get_romw(out.ctypes.data_as(C.c_void_p))
    return out


def table(sed, kind):
    nb = NUMFREQBND if kind < 2 else NUMHEATBIN
    out = np.zeros((nb, NUMTAU + 1))
    lib().orc_get_table(C.c_int(sed), C.c_int(kind), out.ctypes.data_as(C.c_void_p))
    return out


def rec_colion(T):
    out = np.zeros(12)
    lib().orc_rec_colion(C.c_double(T), out.ctypes.data_as(C.c_void_p))
    return out


def set_params(isothermal, temper_val=1e4, clumping=1.0, zred=9.0, H0=0.0, Omega0=0.27, cosmological=True,
               subboxsize=10, max_subbox=1150):
    lib().orc_set_params(C.c_int(int(isothermal)), C.c_double(temper_val), C.c_float(clumping), C.c_double(zred),
                         C.c_double(H0), C.c_double(Omega0), C.c_int(int(cosmological)), C.c_int(subboxsize),
                         C.c_int(max_subbox))


def coolin(n, ne, xh, xhe, T):
    return lib().orc_coolin(n, ne, np.ascontiguousarray(xh, dtype=np.float64), np.ascontiguousarray(xhe, dtype=np.float64), T)


def doric(dt, rhe, rhh, ion15, phi3, fr4, T):
    ion = np.array(ion15, dtype=np.float64)
    lib().orc_doric(C.c_double(dt), C.c_double(rhe), C.c_double(rhh), ion.ctypes.data_as(C.c_void_p),
                    np.asarray(phi3, dtype=np.float64).ctypes.data_as(C.c_void_p),
                    np.asarray(fr4, dtype=np.float64).ctypes.data_as(C.c_void_p), C.c_double(T))
    return ion


def thermal(dt, T, n_e, n, ion15, heat):
    e, a, ns = C.c_double(T), C.c_double(0.0), C.c_int(0)
    lib().orc_thermal(C.c_double(dt), C.byref(e), C.byref(a), C.c_double(n_e), C.c_double(n),
                      np.asarray(ion15, dtype=np.float64).ctypes.data_as(C.c_void_p), C.c_double(heat), C.byref(ns))
    return e.value, a.value, ns.value


def chemistry_batch(dt, ndens, ion15, phi4, T3):
    n = len(ndens)
    ion = np.array(ion15, dtype=np.float64).reshape(n, 15).copy()
    T = np.array(T3, dtype=np.float64).reshape(n, 3).copy()
    nit = np.zeros(n, dtype=np.int32)
    lib().orc_chemistry_batch(C.c_int(n), C.c_double(dt), np.ascontiguousarray(ndens, dtype=np.float64).ctypes.data_as(C.c_void_p),
                              ion.ctypes.data_as(C.c_void_p),
                              np.ascontiguousarray(phi4, dtype=np.float64).ctypes.data_as(C.c_void_p),
                              T.ctypes.data_as(C.c_void_p), nit.ctypes.data_as(C.c_void_p))
    return ion, T, nit


def photoion_rates_batch(col6, vol, nflux3, i_state):
    col6 = np.ascontiguousarray(col6, dtype=np.float64)
    n = col6.shape[0]
    out = np.zeros((n, 6))
    lib().orc_photoion_rates_batch(C.c_int(n), col6.ctypes.data_as(C.c_void_p),
                                   np.ascontiguousarray(vol, dtype=np.float64).ctypes.data_as(C.c_void_p),
                                   np.ascontiguousarray(nflux3, dtype=np.float64).ctypes.data_as(C.c_void_p),
                                   np.ascontiguousarray(i_state, dtype=np.float64).ctypes.data_as(C.c_void_p),
                                   out.ctypes.data_as(C.c_void_p))
    return out


class Grid:
    """Grid-sized oracle state (module-global in the library: one Grid at a time)."""

    def __init__(self, mesh, dr, vol):
        self.mesh = np.asarray(mesh, dtype=np.int32)
        self.N3 = int(np.prod(self.mesh))
        lib().orc_grid_init(self.mesh.ctypes.data_as(C.c_void_p), np.asarray(dr, dtype=np.float64).ctypes.data_as(C.c_void_p),
                            C.c_double(vol))

    def set_state(self, ndens, xh, xhe, temperature_grid=None):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (ndens, xh, xhe)]
        t = None if temperature_grid is None else np.ascontiguousarray(temperature_grid, dtype=np.float32)
        lib().orc_set_state(*[x.ctypes.data_as(C.c_void_p) for x in a], None if t is None else t.ctypes.data_as(C.c_void_p))

    def set_clumping_grid(self, grid):
        g = None if grid is None else np.ascontiguousarray(grid, dtype=np.float32)
        lib().orc_set_clumping_grid(None if g is None else g.ctypes.data_as(C.c_void_p))

    def set_LLS(self, type_of_LLS, coldensh_LLS=0.0, LLS_grid=None):
        g = None if LLS_grid is None else np.ascontiguousarray(LLS_grid, dtype=np.float32)
        lib().orc_set_LLS(C.c_int(type_of_LLS), C.c_double(coldensh_LLS), None if g is None else g.ctypes.data_as(C.c_void_p))

    def set_work_state(self, xh_av, xhe_av, xh_intermed, xhe_intermed):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (xh_av, xhe_av, xh_intermed, xhe_intermed)]
        lib().orc_set_work_state(*[x.ctypes.data_as(C.c_void_p) for x in a])

    def set_rates(self, phih, phihe, phiheat):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (phih, phihe, phiheat)]
        lib().orc_set_rates(*[x.ctypes.data_as(C.c_void_p) for x in a])

    def set_sources(self, srcpos, NormFlux, NormFluxPL=None, NormFluxQPL=None):
        sp = np.ascontiguousarray(srcpos, dtype=np.int32).reshape(-1, 3)
        n = sp.shape[0]
        f = [None if x is None else np.ascontiguousarray(x, dtype=np.float64) for x in (NormFlux, NormFluxPL, NormFluxQPL)]
        lib().orc_set_sources(C.c_int(n), sp.ctypes.data_as(C.c_void_p),
                              *[None if x is None else x.ctypes.data_as(C.c_void_p) for x in f])
        self.NumSrc = n

    def get_state(self):
        m = tuple(self.mesh[::-1])
        xh, xhe, T = np.zeros((2,) + m), np.zeros((3,) + m), np.zeros((3,) + m, dtype=np.float32)
        lib().orc_get_state(xh.ctypes.data_as(C.c_void_p), xhe.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p))
        return xh, xhe, T

    def get_work_state(self):
        m = tuple(self.mesh[::-1])
        a = [np.zeros((2,) + m), np.zeros((3,) + m), np.zeros((2,) + m), np.zeros((3,) + m)]
        lib().orc_get_work_state(*[x.ctypes.data_as(C.c_void_p) for x in a])
        return a

    def get_rates(self):
        m = tuple(self.mesh[::-1])
        a = [np.zeros(m), np.zeros((2,) + m), np.zeros(m)]
        lib().orc_get_rates(*[x.ctypes.data_as(C.c_void_p) for x in a])
        return a

    def set_rates_to_zero(self):
        lib().orc_set_rates_to_zero()

    def pass_all_sources(self, nthreads=1, order=0, rank=0, npr=1):
        nbox = np.zeros(self.NumSrc, dtype=np.int32)
        upd = lib().orc_pass_all_sources(C.c_int(nthreads), C.c_int(order), C.c_int(rank), C.c_int(npr),
                                         nbox.ctypes.data_as(C.c_void_p))
        return upd, nbox, lib().orc_photon_loss(), lib().orc_sum_nbox()

    def global_pass(self, dt, nthreads=1, want_nit=False):
        nit = np.zeros(self.N3, dtype=np.int32) if want_nit else None
        cf = lib().orc_global_pass(C.c_double(dt), C.c_int(nthreads), None if nit is None else nit.ctypes.data_as(C.c_void_p))
        return (cf, nit) if want_nit else cf

    def global_pass_range(self, dt, p0, p1):
        return lib().orc_global_pass_range(C.c_double(dt), C.c_long(p0), C.c_long(p1))

    def evolve3d(self, dt, nthreads=1, order=0):
        stats = np.zeros(5, dtype=np.int64)
        hist = np.zeros(512, dtype=np.int32)
        lib().orc_evolve3d(C.c_double(dt), C.c_int(nthreads), C.c_int(order), stats.ctypes.data_as(C.c_void_p),
                           hist.ctypes.data_as(C.c_void_p))
        return dict(niter=int(stats[0]), conv_flag=int(stats[1]), conv_criterion=int(stats[2]), sum_nbox=int(stats[3]),
                    rt_updates=int(stats[4]), conv_hist=hist[1:int(stats[0]) + 1].copy())

    def total_rates(self, dt, xh_av, xhe_av):
        out = np.zeros(3)
        lib().orc_total_rates(C.c_double(dt), np.ascontiguousarray(xh_av).ctypes.data_as(C.c_void_p),
                              np.ascontiguousarray(xhe_av).ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    def state_sums(self, xh, xhe):
        out = np.zeros(5)
        lib().orc_state_sums(np.ascontiguousarray(xh).ctypes.data_as(C.c_void_p), np.ascontiguousarray(xhe).ctypes.data_as(C.c_void_p),
                             out.ctypes.data_as(C.c_void_p))
        return out


def mrgrnk(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(len(x), dtype=np.int32)
    lib().orc_mrgrnk(C.c_int(len(x)), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n=None):
    """n=None: every CPU this process may run on (a launcher's OMP_NUM_THREADS=1 default is overridden)."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()
