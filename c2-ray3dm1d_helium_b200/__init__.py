"""c2-ray3dm1d_helium_b200 -- B200-native (sm_100a) hot path of C2-Ray H+He: the 3D short-characteristics sweep,
the photo-ionization / heating table lookups and the doric + thermal chemistry, behind the C ABI in
include/c2ray_b200.h.  Import via `import c2ray_b200` (root shim) or importlib (the directory name has a hyphen).

Layout: csrc/ (CUDA kernels + C ABI), capi.py (ctypes binding), evolve.py (mirror of the reference's evolve /
evolve_source / radiation_* / doric interfaces), synth.py (synthetic inputs of the BASELINE configs).
"""
from . import capi, synth  # noqa: F401
from .evolve import (C2Ray, C2RayParameters, balanced_partition, fortran_records_read, fortran_records_write, from_problem,  # noqa: F401
                     read_cooling_tables, source_partition)
