"""ctypes binding of include/c2ray_b200.h (libc2ray_b200.so) -- the same C ABI a Fortran host binds through
iso_c_binding.  Loading fails loudly when the library has not been built; there is no CPU fallback."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, os.environ.get("C2RAY_B200_LIB", "libc2ray_b200.so"))  # env override: tuning builds only
CSRC = os.path.join(HERE, "csrc")

NUMFREQBND, NUMHEATBIN, NUMTAU, MAX_ITER_HIST = 47, 113, 2000, 512

EXPORTS = [
    "c2ray_b200_last_error", "c2ray_b200_init", "c2ray_b200_destroy", "c2ray_b200_set_params",
    "c2ray_b200_set_cooling_tables", "c2ray_b200_rad_ini", "c2ray_b200_upload_tables", "c2ray_b200_download_table",
    "c2ray_b200_set_sources", "c2ray_b200_set_geometry", "c2ray_b200_set_state", "c2ray_b200_get_state",
    "c2ray_b200_get_rates", "c2ray_b200_set_rates", "c2ray_b200_get_work_state", "c2ray_b200_set_work_state",
    "c2ray_b200_snapshot_state", "c2ray_b200_restore_state", "c2ray_b200_evolve3d", "c2ray_b200_evolve3d_host",
    "c2ray_b200_begin_step", "c2ray_b200_set_rates_to_zero", "c2ray_b200_pass_all_sources", "c2ray_b200_do_source",
    "c2ray_b200_global_pass", "c2ray_b200_end_step", "c2ray_b200_state_sums", "c2ray_b200_photoion_rates_batch",
    "c2ray_b200_chemistry_batch", "c2ray_b200_doric_batch", "c2ray_b200_thermal_batch", "c2ray_b200_rec_colion_batch", "c2ray_b200_cinterp_batch",
    "c2ray_b200_comm_unique_id", "c2ray_b200_comm_init", "c2ray_b200_set_rank", "c2ray_b200_rates_device_buffer",
    "c2ray_b200_bench_global_pass", "c2ray_b200_launch_count", "c2ray_b200_sweep_launch_count", "c2ray_b200_measure_fp64", "c2ray_b200_stream",
    "c2ray_b200_timer_start", "c2ray_b200_timer_stop",
    "c2ray_b200_set_dump", "c2ray_b200_write_iteration_dump", "c2ray_b200_read_iteration_dump",
    "c2ray_b200_write_stream2", "c2ray_b200_write_stream3", "c2ray_b200_fortran_records_write",
    "c2ray_b200_fortran_records_read", "c2ray_b200_set_clumping_grid", "c2ray_b200_set_LLS",
    "c2ray_b200_set_source_schedule", "c2ray_b200_balanced_partition", "c2ray_b200_my_sources", "c2ray_b200_mrgrnk",
]


class Params(C.Structure):
    _fields_ = [("isothermal", C.c_int32), ("cosmological", C.c_int32), ("subboxsize", C.c_int32),
                ("max_subbox", C.c_int32), ("temper_val", C.c_double), ("H0", C.c_double), ("Omega0", C.c_double),
                ("clumping", C.c_float), ("max_slots", C.c_int32), ("deterministic", C.c_int32)]


class SedParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("T_eff", "S_star", "pl_index", "pl_minfreq", "pl_maxfreq", "pl_S_star",
                                          "qpl_index", "qpl_minfreq", "qpl_maxfreq", "qpl_S_star")]


class SedTables(C.Structure):
    _fields_ = [("photo_thick", C.c_void_p), ("photo_thin", C.c_void_p), ("heat_thick", C.c_void_p),
                ("heat_thin", C.c_void_p), ("freqbnd_lower", C.c_int32), ("freqbnd_upper", C.c_int32),
                ("S_star", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("niter", C.c_int32), ("conv_flag", C.c_int32), ("conv_criterion", C.c_int32), ("nit_max", C.c_int32),
                ("sum_nbox_all", C.c_int64), ("rt_updates", C.c_int64), ("chem_cells", C.c_int64),
                ("nit_total", C.c_int64), ("photon_loss_all", C.c_double), ("ms_sweep", C.c_double),
                ("ms_chem", C.c_double), ("ms_allreduce", C.c_double), ("ms_total", C.c_double),
                ("sums_before", C.c_double * 5), ("sums_after", C.c_double * 5), ("totrec", C.c_double),
                ("totcollisions", C.c_double), ("recomions", C.c_double), ("total_ion", C.c_double), ("totalsrc", C.c_double),
                ("photcons", C.c_double), ("conv_hist", C.c_int32 * MAX_ITER_HIST)]


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libc2ray_b200.so (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC, "-s"] if not verbose else ["make", "-C", CSRC])
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                           "The C2-Ray B200 hot path has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.c2ray_b200_last_error.restype = C.c_char_p
    L.c2ray_b200_launch_count.restype = C.c_int64
    L.c2ray_b200_launch_count.argtypes = [C.c_void_p]
    if hasattr(L, "c2ray_b200_sweep_launch_count"):
        L.c2ray_b200_sweep_launch_count.restype = C.c_int64
        L.c2ray_b200_sweep_launch_count.argtypes = [C.c_void_p]
    for name in EXPORTS:
        try:
            fn = getattr(L, name)
        except AttributeError:
            if "C2RAY_B200_LIB" in os.environ:  # an older tuning build loaded for an A/B timing
                continue
            raise
        if name not in ("c2ray_b200_last_error", "c2ray_b200_launch_count", "c2ray_b200_sweep_launch_count"):
            fn.restype = C.c_int
    _lib = L
    return L


class C2RayError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise C2RayError(f"libc2ray_b200 error {rc}: {load().c2ray_b200_last_error().decode()}")
