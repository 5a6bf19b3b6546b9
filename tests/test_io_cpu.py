"""CPU checks of the Fortran unformatted-record layer behind the iteration dumps and output streams
(evolve.F90:233-367, output.F90:249-379): what the library writes must be what gfortran/ifort would read.  The
independent reader is scipy.io.FortranFile; the subrecord convention (records above 2^31-9 bytes) is checked against
the marker rules with a small subrecord limit."""
import struct

import numpy as np
import pytest
from scipy.io import FortranFile

import c2ray_b200
from c2ray_b200 import capi


def test_records_are_readable_by_an_independent_fortran_reader(tmp_path):
    rng = np.random.default_rng(1)
    mesh = np.array([6, 5, 4], dtype=np.int32)
    niter = np.array([7], dtype=np.int32)
    loss = rng.random(47)
    grid = rng.random(6 * 5 * 4)
    t32 = rng.random(3 * 6 * 5 * 4).astype(np.float32)
    path = tmp_path / "iterdump1.bin"
    c2ray_b200.fortran_records_write(path, [niter, loss, grid, t32, mesh])
    f = FortranFile(path, "r")
    assert f.read_ints(np.int32)[0] == 7
    assert np.array_equal(f.read_reals(np.float64), loss)
    assert np.array_equal(f.read_reals(np.float64), grid)
    assert np.array_equal(f.read_reals(np.float32), t32)
    assert np.array_equal(f.read_ints(np.int32), mesh)
    f.close()
    # and back through the library's own reader
    back = c2ray_b200.fortran_records_read(path, [(np.int32, 1), (np.float64, 47), (np.float64, grid.size),
                                                  (np.float32, t32.size), (np.int32, 3)])
    assert back[0][0] == 7 and np.array_equal(back[2], grid) and np.array_equal(back[3], t32)


def test_library_reads_what_a_fortran_writer_wrote(tmp_path):
    path = tmp_path / "xfrac3d_9.000.bin"
    a = np.arange(24, dtype=np.float64) / 7
    f = FortranFile(path, "w")
    f.write_record(np.array([2, 3, 4], dtype=np.int32))
    f.write_record(a)
    f.close()
    m, b = c2ray_b200.fortran_records_read(path, [(np.int32, 3), (np.float64, 24)])
    assert list(m) == [2, 3, 4] and np.array_equal(a, b)


def test_subrecords_follow_the_compilers_convention(tmp_path):
    """gfortran/ifort: leading marker negative when another subrecord follows, trailing marker negative when one
    preceded it; |marker| = bytes of the subrecord."""
    path = tmp_path / "big.bin"
    a = np.arange(25, dtype=np.float64)           # 200 bytes, subrecords of at most 64 -> 64+64+64+8
    tail = np.array([42], dtype=np.int32)
    c2ray_b200.fortran_records_write(path, [a, tail], max_subrecord=64)
    raw = open(path, "rb").read()
    pos, payload, markers = 0, b"", []
    while True:
        lead = struct.unpack_from("<i", raw, pos)[0]
        n = abs(lead)
        payload += raw[pos + 4:pos + 4 + n]
        trail = struct.unpack_from("<i", raw, pos + 4 + n)[0]
        markers.append((lead, trail))
        pos += 8 + n
        if lead > 0:
            break
    assert markers == [(-64, 64), (-64, -64), (-64, -64), (8, -8)]
    assert np.array_equal(np.frombuffer(payload, dtype=np.float64), a)
    assert struct.unpack_from("<iii", raw, pos) == (4, 42, 4) and pos + 12 == len(raw)
    b, t = c2ray_b200.fortran_records_read(path, [(np.float64, 25), (np.int32, 1)])
    assert np.array_equal(a, b) and t[0] == 42
    # exactly one full subrecord: no continuation
    c2ray_b200.fortran_records_write(path, [a[:8]], max_subrecord=64)
    assert struct.unpack_from("<i", open(path, "rb").read(), 0)[0] == 64


def test_length_mismatch_and_missing_files_are_errors(tmp_path):
    path = tmp_path / "d.bin"
    c2ray_b200.fortran_records_write(path, [np.zeros(10), np.zeros(3, dtype=np.int32)])
    with pytest.raises(capi.C2RayError):
        c2ray_b200.fortran_records_read(path, [(np.float64, 9), (np.int32, 3)])     # record longer than expected
    with pytest.raises(capi.C2RayError):
        c2ray_b200.fortran_records_read(path, [(np.float64, 11), (np.int32, 3)])    # record shorter than expected
    with pytest.raises(capi.C2RayError):
        c2ray_b200.fortran_records_read(path, [(np.float64, 10), (np.int32, 3), (np.int32, 1)])  # one record too many
    with pytest.raises(capi.C2RayError):
        c2ray_b200.fortran_records_read(tmp_path / "absent.bin", [(np.int32, 1)])
    with pytest.raises(capi.C2RayError):
        c2ray_b200.fortran_records_write(tmp_path / "no_such_dir" / "x.bin", [np.zeros(1)])
    # empty record
    c2ray_b200.fortran_records_write(path, [np.zeros(0), np.ones(2)])
    assert open(path, "rb").read()[:8] == b"\0" * 8
    e, o = c2ray_b200.fortran_records_read(path, [(np.float64, 0), (np.float64, 2)])
    assert e.size == 0 and list(o) == [1.0, 1.0]
