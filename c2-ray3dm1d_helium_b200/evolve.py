"""Host-side mirror of the reference's interface for the hot path, over the C ABI (include/c2ray_b200.h).

Names and argument meaning follow the Fortran modules:
  evolve (files_for_3D/evolve.F90): evolve3D(time,dt,restart) :78, pass_all_sources :385, global_pass :435,
      set_rates_to_zero :371
  evolve_source.F90: do_source(dt,ns1,niter) :66
  radiation_tables.f90: rad_ini() :141 ; radiation_photoionrates.f90: photoion_rates :108
  cooling_h.f90: setup_cool() :76 ; evolve_point.F90: do_chemistry :444 ; cgsconstants.f90: ini_rec_colion_factors :140
  c2ray_parameters.f90 -> C2RayParameters
Arrays use the Fortran memory layout: numpy C-order arrays indexed [component, k, j, i]; srcpos is (NumSrc,3) 1-based.
Everything computes on the GPU through libc2ray_b200.so; there is no CPU path in this package.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import capi

NUMTAU, NUMFREQBND, NUMHEATBIN = capi.NUMTAU, capi.NUMFREQBND, capi.NUMHEATBIN


@dataclass
class C2RayParameters:
    """Run-time values of c2ray_parameters.f90:26-89 (+ the material/cosmology scalars the path reads)."""
    isothermal: bool = False
    cosmological: bool = True
    subboxsize: int = 10
    max_subbox: int = 1150
    temper_val: float = 1.0e4
    H0: float = 0.0
    Omega0: float = 0.27
    clumping: float = 1.0
    max_slots: int = 0
    deterministic: bool = False

    def to_c(self):
        return capi.Params(int(self.isothermal), int(self.cosmological), int(self.subboxsize), int(self.max_subbox),
                           float(self.temper_val), float(self.H0), float(self.Omega0), float(self.clumping),
                           int(self.max_slots), int(self.deterministic))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def read_cooling_tables(path=None):
    """The five curves cooling_h.f90:83-149 reads, from data/cooling_h_he.tab: returns logT[801], logLambda[5,801]."""
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "cooling_h_he.tab")
    a = np.loadtxt(path)
    if a.shape != (801, 6):
        raise ValueError("cooling table must have 801 rows of logT + 5 curves")
    return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1:].T)


class C2Ray:
    """One rank's view of the hot path: owns a c2ray_ctx (device-resident grids) and mirrors the module procedures."""

    def __init__(self, mesh, params=None, device=-1):
        self.lib = capi.load()
        self.params = params or C2RayParameters()
        self.mesh = np.asarray(mesh, dtype=np.int32).copy()
        self.N3 = int(np.prod(self.mesh.astype(np.int64)))
        self._shape = tuple(int(x) for x in self.mesh[::-1])
        self.ctx = C.c_void_p()
        cp = self.params.to_c()
        capi.check(self.lib.c2ray_b200_init(C.byref(cp), _p(self.mesh), C.c_int32(device), C.byref(self.ctx)))
        self.NumSrc = 0
        self.last_stats = None

    # -- life cycle -------------------------------------------------------------------------------------
    def close(self):
        if self.ctx:
            self.lib.c2ray_b200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params):
        self.params = params
        cp = params.to_c()
        capi.check(self.lib.c2ray_b200_set_params(self.ctx, C.byref(cp)))

    # -- setup (cooling_h.f90 setup_cool, radiation_tables.f90 rad_ini, sourceprops, grid) -----------------
    def setup_cool(self, logT=None, logLambda=None):
        if logT is None:
            logT, logLambda = read_cooling_tables()
        logT, logLambda = _f64(logT), _f64(logLambda)
        assert logT.shape == (801,) and logLambda.shape == (5, 801)
        capi.check(self.lib.c2ray_b200_set_cooling_tables(self.ctx, _p(logT), _p(logLambda)))

    def rad_ini(self, T_eff=5.0e4, S_star=1e48, pl=None, qpl=None):
        z = dict(index=1.0, minfreq=1.0, maxfreq=2.0, S_star=0.0)
        p, q = pl or z, qpl or z
        sp = capi.SedParams(T_eff, S_star, p["index"], p["minfreq"], p["maxfreq"], p["S_star"], q["index"], q["minfreq"],
                            q["maxfreq"], q["S_star"])
        capi.check(self.lib.c2ray_b200_rad_ini(self.ctx, C.byref(sp)))

    def upload_tables(self, sed, photo_thick, photo_thin, heat_thick, heat_thin, lower, upper, S_star):
        """Tables as the host's radiation_tables module holds them: arrays [band, 0:NumTau]."""
        if photo_thick is None:
            t = capi.SedTables(None, None, None, None, 1, 0, 0.0)
            capi.check(self.lib.c2ray_b200_upload_tables(self.ctx, C.c_int32(sed), C.byref(t)))
            return
        a = [_f64(photo_thick), _f64(photo_thin), None if heat_thick is None else _f64(heat_thick),
             None if heat_thin is None else _f64(heat_thin)]
        assert a[0].shape == (NUMFREQBND, NUMTAU + 1) and a[1].shape == a[0].shape
        t = capi.SedTables(a[0].ctypes.data, a[1].ctypes.data, None if a[2] is None else a[2].ctypes.data,
                           None if a[3] is None else a[3].ctypes.data, int(lower), int(upper), float(S_star))
        capi.check(self.lib.c2ray_b200_upload_tables(self.ctx, C.c_int32(sed), C.byref(t)))

    def download_table(self, sed, kind):
        nb = NUMFREQBND if kind < 2 else NUMHEATBIN
        out = np.zeros((nb, NUMTAU + 1))
        lo, hi, S = C.c_int32(), C.c_int32(), C.c_double()
        capi.check(self.lib.c2ray_b200_download_table(self.ctx, C.c_int32(sed), C.c_int32(kind), _p(out), C.byref(lo),
                                                      C.byref(hi), C.byref(S)))
        return out, lo.value, hi.value, S.value

    def sed_limits(self, sed):
        lo, hi, S = C.c_int32(), C.c_int32(), C.c_double()
        capi.check(self.lib.c2ray_b200_download_table(self.ctx, C.c_int32(sed), C.c_int32(0), None, C.byref(lo),
                                                      C.byref(hi), C.byref(S)))
        return lo.value, hi.value, S.value

    def set_sources(self, srcpos, NormFlux, NormFluxPL=None, NormFluxQPL=None):
        sp = np.ascontiguousarray(srcpos, dtype=np.int32).reshape(-1, 3)
        n = sp.shape[0]
        f = [None if x is None else _f64(x) for x in (NormFlux, NormFluxPL, NormFluxQPL)]
        capi.check(self.lib.c2ray_b200_set_sources(self.ctx, C.c_int32(n), _p(sp), *[_p(x) for x in f]))
        self.NumSrc = n

    def set_geometry(self, dr, vol, zred):
        dr = _f64(dr)
        capi.check(self.lib.c2ray_b200_set_geometry(self.ctx, _p(dr), C.c_double(vol), C.c_double(zred)))

    def set_clumping_grid(self, clumping_grid):
        """type_of_clumping == 5: material's clumping_grid(i,j,k) (real); None returns to the scalar clumping."""
        g = None if clumping_grid is None else np.ascontiguousarray(clumping_grid, dtype=np.float32)
        assert g is None or g.size == self.N3
        capi.check(self.lib.c2ray_b200_set_clumping_grid(self.ctx, _p(g)))

    def set_LLS(self, type_of_LLS, coldensh_LLS=0.0, LLS_grid=None):
        """use_LLS: 0 off, 1 one column density per cell (coldensh_LLS), 2 position dependent (LLS_grid, real)."""
        g = None if LLS_grid is None else np.ascontiguousarray(LLS_grid, dtype=np.float32)
        assert g is None or g.size == self.N3
        capi.check(self.lib.c2ray_b200_set_LLS(self.ctx, C.c_int32(type_of_LLS), C.c_double(coldensh_LLS), _p(g)))

    # -- material state -----------------------------------------------------------------------------------
    def set_state(self, ndens, xh, xhe, temperature_grid=None):
        a = [_f64(ndens), _f64(xh), _f64(xhe)]
        assert a[0].size == self.N3 and a[1].size == 2 * self.N3 and a[2].size == 3 * self.N3
        t = None if temperature_grid is None else np.ascontiguousarray(temperature_grid, dtype=np.float32)
        capi.check(self.lib.c2ray_b200_set_state(self.ctx, _p(a[0]), _p(a[1]), _p(a[2]), _p(t)))

    def get_state(self):
        xh, xhe = np.zeros((2,) + self._shape), np.zeros((3,) + self._shape)
        T = np.zeros((3,) + self._shape, dtype=np.float32)
        capi.check(self.lib.c2ray_b200_get_state(self.ctx, _p(xh), _p(xhe), _p(T)))
        return xh, xhe, T

    def get_rates(self):
        a = [np.zeros(self._shape), np.zeros((2,) + self._shape), np.zeros(self._shape)]
        capi.check(self.lib.c2ray_b200_get_rates(self.ctx, *[_p(x) for x in a]))
        return a

    def set_rates(self, phih, phihe, phiheat=None):
        a = [_f64(phih), _f64(phihe), None if phiheat is None else _f64(phiheat)]
        capi.check(self.lib.c2ray_b200_set_rates(self.ctx, *[_p(x) for x in a]))

    def get_work_state(self):
        a = [np.zeros((2,) + self._shape), np.zeros((3,) + self._shape), np.zeros((2,) + self._shape),
             np.zeros((3,) + self._shape)]
        capi.check(self.lib.c2ray_b200_get_work_state(self.ctx, *[_p(x) for x in a]))
        return a

    def set_work_state(self, xh_av, xhe_av, xh_intermed, xhe_intermed):
        a = [_f64(x) for x in (xh_av, xhe_av, xh_intermed, xhe_intermed)]
        capi.check(self.lib.c2ray_b200_set_work_state(self.ctx, *[_p(x) for x in a]))

    def snapshot_state(self):
        capi.check(self.lib.c2ray_b200_snapshot_state(self.ctx))

    def restore_state(self):
        capi.check(self.lib.c2ray_b200_restore_state(self.ctx))

    # -- the hot path ---------------------------------------------------------------------------------------
    def evolve3D(self, time, dt, restart=0):
        """evolve.F90:78 on the device-resident state."""
        st = capi.Stats()
        capi.check(self.lib.c2ray_b200_evolve3d(self.ctx, C.c_double(time), C.c_double(dt), C.c_int32(restart), C.byref(st)))
        self.last_stats = self._stats(st)
        return self.last_stats

    def evolve3D_host(self, time, dt, restart, ndens, xh, xhe, temperature_grid=None):
        """The Fortran-facing call: host arrays in, evolve3D, host arrays out (xh, xhe, temperature_grid updated in place)."""
        st = capi.Stats()
        ndens = _f64(ndens)
        assert xh.dtype == np.float64 and xhe.dtype == np.float64 and xh.flags.c_contiguous and xhe.flags.c_contiguous
        capi.check(self.lib.c2ray_b200_evolve3d_host(self.ctx, C.c_double(time), C.c_double(dt), C.c_int32(restart),
                                                     _p(ndens), _p(xh), _p(xhe), _p(temperature_grid), C.byref(st)))
        self.last_stats = self._stats(st)
        return self.last_stats

    @staticmethod
    def _stats(st):
        d = {k: getattr(st, k) for k in ("niter", "conv_flag", "conv_criterion", "nit_max", "sum_nbox_all", "rt_updates",
                                         "chem_cells", "nit_total", "photon_loss_all", "ms_sweep", "ms_chem", "ms_allreduce",
                                         "ms_total", "totrec", "totcollisions", "recomions", "total_ion", "totalsrc", "photcons")}
        d["sums_before"] = np.array(st.sums_before[:])
        d["sums_after"] = np.array(st.sums_after[:])
        d["conv_hist"] = np.array(st.conv_hist[:min(st.niter, capi.MAX_ITER_HIST)])
        return d

    def begin_step(self):
        capi.check(self.lib.c2ray_b200_begin_step(self.ctx))

    def end_step(self):
        capi.check(self.lib.c2ray_b200_end_step(self.ctx))

    def set_rates_to_zero(self):
        capi.check(self.lib.c2ray_b200_set_rates_to_zero(self.ctx))

    def pass_all_sources(self, niter, dt):
        upd = C.c_int64()
        capi.check(self.lib.c2ray_b200_pass_all_sources(self.ctx, C.c_double(dt), C.c_int32(niter), C.byref(upd)))
        return upd.value

    def do_source(self, dt, ns1, niter):
        nbox, loss = C.c_int32(), C.c_double()
        capi.check(self.lib.c2ray_b200_do_source(self.ctx, C.c_double(dt), C.c_int32(ns1), C.c_int32(niter), C.byref(nbox),
                                                 C.byref(loss)))
        return nbox.value, loss.value

    def global_pass(self, dt, want_nit=False):
        cf = C.c_int32()
        nit = np.zeros(self._shape, dtype=np.int32) if want_nit else None
        capi.check(self.lib.c2ray_b200_global_pass(self.ctx, C.c_double(dt), C.byref(cf), _p(nit)))
        return (cf.value, nit) if want_nit else cf.value

    def state_sums(self, which=0):
        out = np.zeros(5)
        capi.check(self.lib.c2ray_b200_state_sums(self.ctx, C.c_int32(which), _p(out)))
        return out

    # -- iteration dumps, restart, output streams (evolve.F90:233-367, output.F90:249-379) ------------------------
    def set_dump(self, dump_dir, interval_s=15.0 * 60.0):
        """dump_dir: where evolve3D writes iterdump1/2.bin (when interval_s >= 0) and where restart=1|2|3 reads."""
        capi.check(self.lib.c2ray_b200_set_dump(self.ctx, None if dump_dir is None else os.fsencode(dump_dir),
                                                C.c_double(interval_s)))

    def write_iteration_dump(self, path, niter):
        capi.check(self.lib.c2ray_b200_write_iteration_dump(self.ctx, os.fsencode(path), C.c_int32(niter)))

    def start_from_dump(self, path):
        niter = C.c_int32()
        capi.check(self.lib.c2ray_b200_read_iteration_dump(self.ctx, os.fsencode(path), C.byref(niter)))
        return niter.value

    def write_stream2(self, results_dir, zred_now):
        capi.check(self.lib.c2ray_b200_write_stream2(self.ctx, os.fsencode(results_dir), C.c_double(zred_now)))

    def write_stream3(self, results_dir, zred_now):
        capi.check(self.lib.c2ray_b200_write_stream3(self.ctx, os.fsencode(results_dir), C.c_double(zred_now)))

    # -- parity hooks -----------------------------------------------------------------------------------------
    def photoion_rates(self, col6, vol, nflux3, i_state):
        col6 = _f64(col6).reshape(-1, 6)
        n = col6.shape[0]
        out = np.zeros((n, 6))
        capi.check(self.lib.c2ray_b200_photoion_rates_batch(self.ctx, C.c_int32(n), _p(col6), _p(_f64(vol)),
                                                            _p(_f64(nflux3)), _p(_f64(i_state)), _p(out)))
        return out

    def do_chemistry(self, dt, ndens, ion15, phi4, T3):
        n = len(ndens)
        ion = np.array(ion15, dtype=np.float64).reshape(n, 15).copy()
        T = np.array(T3, dtype=np.float64).reshape(n, 3).copy()
        nit = np.zeros(n, dtype=np.int32)
        capi.check(self.lib.c2ray_b200_chemistry_batch(self.ctx, C.c_int32(n), C.c_double(dt), _p(_f64(ndens)), _p(ion),
                                                       _p(_f64(phi4)), _p(T), _p(nit)))
        return ion, T, nit

    def doric(self, dt, rhe, ion15, phi3, fr4, T):
        """doric.f90:35 for n states (coefficients at T per state); returns the updated ion15[n][15]."""
        rhe = _f64(np.atleast_1d(rhe)); n = len(rhe)
        ion = np.array(ion15, dtype=np.float64).reshape(n, 15).copy()
        capi.check(self.lib.c2ray_b200_doric_batch(self.ctx, C.c_int32(n), C.c_double(dt), _p(rhe), _p(ion),
                                                   _p(_f64(np.reshape(phi3, (n, 3)))), _p(_f64(np.reshape(fr4, (n, 4)))),
                                                   _p(_f64(np.atleast_1d(T)))))
        return ion

    def thermal(self, dt, end_temper, avg_temper, ndens_electron, ndens_atom, ion15, heat):
        """thermal.f90:22 for n states; returns (end_temper, avg_temper, sub-steps)."""
        e = np.array(np.atleast_1d(end_temper), dtype=np.float64); n = len(e)
        a = np.array(np.atleast_1d(avg_temper), dtype=np.float64)
        ns = np.zeros(n, dtype=np.int32)
        capi.check(self.lib.c2ray_b200_thermal_batch(self.ctx, C.c_int32(n), C.c_double(dt), _p(e), _p(a),
                                                     _p(_f64(np.atleast_1d(ndens_electron))), _p(_f64(np.atleast_1d(ndens_atom))),
                                                     _p(_f64(np.reshape(ion15, (n, 15)))), _p(_f64(np.atleast_1d(heat))), _p(ns)))
        return e, a, ns

    def ini_rec_colion_factors(self, T):
        T = _f64(np.atleast_1d(T))
        out = np.zeros((len(T), 12))
        capi.check(self.lib.c2ray_b200_rec_colion_batch(self.ctx, C.c_int32(len(T)), _p(T), _p(out)))
        return out

    def cinterp(self, pos, srcpos, coldensh_out, coldenshe_out):
        pos = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
        sp = np.ascontiguousarray(srcpos, dtype=np.int32)
        out = np.zeros((pos.shape[0], 4))
        capi.check(self.lib.c2ray_b200_cinterp_batch(self.ctx, C.c_int32(pos.shape[0]), _p(pos), _p(sp),
                                                     _p(_f64(coldensh_out)), _p(_f64(coldenshe_out)), _p(out)))
        return out

    def mrgrnk(self, xvalt):
        """mrgrnk.f90: rank (1-based argsort, stable) of a real(si) array, on the device."""
        x = np.ascontiguousarray(xvalt, dtype=np.float32)
        out = np.zeros(len(x), dtype=np.int32)
        capi.check(self.lib.c2ray_b200_mrgrnk(self.ctx, C.c_int32(len(x)), _p(x), _p(out)))
        return out

    # -- multi-GPU ----------------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        capi.check(capi.load().c2ray_b200_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, uid, rank, npr):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        capi.check(self.lib.c2ray_b200_comm_init(self.ctx, buf, C.c_int32(rank), C.c_int32(npr)))

    def set_source_schedule(self, mode):
        """0: do_grid_static round robin (master_slave.F90:85); 1: balanced by the last pass's sub-box counts."""
        capi.check(self.lib.c2ray_b200_set_source_schedule(self.ctx, C.c_int32(mode)))

    def my_sources(self):
        n = C.c_int32()
        capi.check(self.lib.c2ray_b200_my_sources(self.ctx, None, C.c_int32(0), C.byref(n)))
        ids = np.zeros(max(n.value, 1), dtype=np.int32)
        capi.check(self.lib.c2ray_b200_my_sources(self.ctx, _p(ids), C.c_int32(n.value), C.byref(n)))
        return ids[:n.value]

    def set_rank(self, rank, npr):
        capi.check(self.lib.c2ray_b200_set_rank(self.ctx, C.c_int32(rank), C.c_int32(npr)))

    # -- measurement ----------------------------------------------------------------------------------------------
    def bench_global_pass(self, dt, reps=1):
        ms, cf = C.c_double(), C.c_int32()
        capi.check(self.lib.c2ray_b200_bench_global_pass(self.ctx, C.c_double(dt), C.c_int32(reps), C.byref(ms), C.byref(cf)))
        return ms.value, cf.value

    def timer_start(self):
        capi.check(self.lib.c2ray_b200_timer_start(self.ctx))

    def timer_stop(self):
        ms = C.c_double()
        capi.check(self.lib.c2ray_b200_timer_stop(self.ctx, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.c2ray_b200_launch_count(self.ctx))

    def sweep_launch_count(self):
        return int(self.lib.c2ray_b200_sweep_launch_count(self.ctx))

    def measure_fp64(self):
        t = C.c_double()
        capi.check(self.lib.c2ray_b200_measure_fp64(self.ctx, C.byref(t)))
        return t.value


def fortran_records_write(path, arrays, max_subrecord=0):
    """Write numpy arrays as Fortran unformatted sequential records (one record each) through the library's record
    layer -- no device needed."""
    arrays = [np.ascontiguousarray(a) for a in arrays]
    n = len(arrays)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
    sizes = (C.c_int64 * n)(*[a.nbytes for a in arrays])
    capi.check(capi.load().c2ray_b200_fortran_records_write(os.fsencode(path), C.c_int32(n), ptrs, sizes,
                                                            C.c_int64(max_subrecord)))


def fortran_records_read(path, specs):
    """Read records of known (dtype, count) back: returns a list of 1-D arrays; raises on any length mismatch."""
    arrays = [np.zeros(cnt, dtype=dt) for dt, cnt in specs]
    n = len(arrays)
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
    sizes = (C.c_int64 * n)(*[a.nbytes for a in arrays])
    capi.check(capi.load().c2ray_b200_fortran_records_read(os.fsencode(path), C.c_int32(n), ptrs, sizes))
    return arrays


def balanced_partition(cost, npr):
    """owner[i] = rank that traces source i under the balanced schedule (longest processing time first)."""
    cost = np.ascontiguousarray(cost, dtype=np.int64)
    owner = np.zeros(len(cost), dtype=np.int32)
    capi.check(capi.load().c2ray_b200_balanced_partition(C.c_int32(len(cost)), _p(cost), C.c_int32(npr), _p(owner)))
    return owner


def source_partition(NumSrc, rank, npr):
    """master_slave.F90:85 do_grid_static: ns1 = 1+rank, NumSrc, npr (1-based source numbers of this rank)."""
    return list(range(1 + rank, NumSrc + 1, npr))


def from_problem(p, device=-1, deterministic=False, max_slots=0, tables=None):
    """Build a C2Ray context from a synth.make_problem dict.  tables: optional dict sed -> tuple for upload_tables
    (the Fortran host's own rad_ini output); otherwise the device rad_ini runs."""
    par = C2RayParameters(isothermal=p["isothermal"], cosmological=p["cosmological"], subboxsize=p["subboxsize"],
                          max_subbox=p["max_subbox"], temper_val=p["temper_val"], H0=p["H0"], Omega0=p["Omega0"],
                          clumping=p["clumping"], deterministic=deterministic, max_slots=max_slots)
    c = C2Ray(p["mesh"], par, device=device)
    c.setup_cool()
    c.set_geometry(p["dr"], p["vol"], p["zred"])
    if tables is None:
        c.rad_ini(p["T_eff"], p["S_star"], pl=p.get("pl"), qpl=p.get("qpl"))
    else:
        for sed in range(3):
            t = tables.get(sed)
            if t is None:
                c.upload_tables(sed, None, None, None, None, 1, 0, 0.0)
            else:
                c.upload_tables(sed, *t)
    c.set_sources(p["srcpos"], p["NormFlux"], p.get("NormFluxPL"), p.get("NormFluxQPL"))
    c.set_state(p["ndens"], p["xh"], p["xhe"], p["temperature_grid"])
    return c
