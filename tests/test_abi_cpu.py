"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/c2ray_b200.h declares, its
structs have the layout the ctypes (and iso_c_binding) mirrors assume, and without a GPU the product fails loudly
instead of falling back to any CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import c2ray_b200
from c2ray_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "c2ray_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(c2ray_b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = declared_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(capi.EXPORTS) == names  # the Python binding covers the whole header, nothing more


def test_header_cites_the_reference_for_each_entry_point():
    src = open(HEADER).read()
    for ref in ("evolve.F90:78", "evolve_source.F90:66", "evolve.F90:435", "radiation_tables.f90:141", "radiation_photoionrates.f90:108",
                "evolve_point.F90:444", "cooling_h.f90:76", "evolve.F90:505", "column_density.f90:28", "cgsconstants.f90:140",
                "evolve.F90:233", "evolve.F90:279", "output.F90:249", "output.F90:312", "master_slave.F90:124", "mrgrnk.f90",
                "evolve_point.F90:484", "evolve_point.F90:177"):
        assert ref in src, ref


def test_struct_layouts_match_the_header(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "c2ray_b200.h"\nint main(){'
                    'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(c2ray_params), offsetof(c2ray_params, temper_val),'
                    'offsetof(c2ray_params, clumping), sizeof(c2ray_sed_params), sizeof(c2ray_sed_tables), sizeof(c2ray_stats),'
                    'offsetof(c2ray_stats, photon_loss_all), offsetof(c2ray_stats, conv_hist)); return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(capi.Params), capi.Params.temper_val.offset, capi.Params.clumping.offset, C.sizeof(capi.SedParams),
            C.sizeof(capi.SedTables), C.sizeof(capi.Stats), capi.Stats.photon_loss_all.offset, capi.Stats.conv_hist.offset]
    assert got == want


def test_no_gpu_means_a_loud_failure_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.C2RayError) as e:
        c2ray_b200.C2Ray([8, 8, 8])
    assert "-2" in str(e.value) and "no CUDA device" in str(e.value)


def test_bad_arguments_are_rejected():
    lib = capi.load()
    ctx = C.c_void_p()
    par = c2ray_b200.C2RayParameters().to_c()
    mesh = np.array([1, 8, 8], dtype=np.int32)
    assert lib.c2ray_b200_init(C.byref(par), mesh.ctypes.data_as(C.c_void_p), 0, C.byref(ctx)) == -1
    assert lib.c2ray_b200_init(None, mesh.ctypes.data_as(C.c_void_p), 0, C.byref(ctx)) == -1
    assert lib.c2ray_b200_evolve3d(None, C.c_double(0), C.c_double(1), 0, None) == -1
    assert b"null" in lib.c2ray_b200_last_error()


def test_product_never_touches_the_oracle():
    """Nothing under the package (sources, binding, kernels) may include, import or link anything from oracle/."""
    pkg = os.path.join(ROOT, "c2-ray3dm1d_helium_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), f
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_source_partition_is_do_grid_static():
    # master_slave.F90:85  do ns1=1+rank,NumSrc,npr
    assert c2ray_b200.source_partition(10, 0, 4) == [1, 5, 9]
    assert c2ray_b200.source_partition(10, 3, 4) == [4, 8]
    assert c2ray_b200.source_partition(2, 3, 4) == []
    allsrc = sorted(sum((c2ray_b200.source_partition(1000, r, 8) for r in range(8)), []))
    assert allsrc == list(range(1, 1001))


def test_balanced_partition_is_a_deterministic_lpt_deal():
    """The stand-in for the master/slave hand-out (master_slave.F90:124-326): every source exactly once, heavy sources
    spread first, never worse than the static round robin on a skewed cost list, identical on every call."""
    rng = np.random.default_rng(3)
    cost = (rng.pareto(1.5, 500) * 1000 + 100).astype(np.int64)
    for npr in (1, 2, 3, 8):
        owner = c2ray_b200.balanced_partition(cost, npr)
        assert owner.min() >= 0 and owner.max() < npr and np.array_equal(owner, c2ray_b200.balanced_partition(cost, npr))
        load = np.bincount(owner, weights=cost, minlength=npr)
        static = np.array([cost[r::npr].sum() for r in range(npr)])
        assert load.max() <= static.max()
        assert load.max() <= load.mean() + cost.max()          # the LPT bound
    # equal costs reproduce do_grid_static's round robin
    assert list(c2ray_b200.balanced_partition(np.full(10, 7), 4)) == [0, 1, 2, 3, 0, 1, 2, 3, 0, 1]
    assert c2ray_b200.balanced_partition(np.zeros(0, dtype=np.int64), 4).size == 0
