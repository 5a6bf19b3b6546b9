"""BASELINE config 5: chemistry-only (doric + thermal) global pass over a slab of synthetic cells.
usage: bench_chem.py [n (slab is n^3 cells)] [reps] [iso]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import c2ray_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
iso = len(sys.argv) > 3 and sys.argv[3] == "iso"
q = c2ray_b200.synth.make_chemistry_problem(n ** 3, isothermal=iso)
par = c2ray_b200.C2RayParameters(isothermal=iso, H0=q["H0"], Omega0=q["Omega0"])
c = c2ray_b200.C2Ray([n, n, n], par, device=0)
c.setup_cool()
c.set_geometry([1e22] * 3, 1e66, q["zred"])
c.set_state(q["ndens"], q["xh"], q["xhe"], q["temperature_grid"])
c.snapshot_state()
c.set_rates(q["phih"], q["phihe"], q["phiheat"])
ms, cf = c.bench_global_pass(q["dt"], 2)
ms, cf = c.bench_global_pass(q["dt"], reps)
cells = n ** 3
B = 200 if iso else 224
print(f"chemistry {'isothermal' if iso else 'thermal'} {n}^3 cells: {ms:.3f} ms/pass, {cells/ms/1e6:.2f} G cells/s, "
      f"{B*cells/ms/1e6:.0f} GB/s algorithmic ({B} B/cell) = {B*cells/ms/1e6/6553:.3f} of 6553 GB/s ; conv_flag={cf}")
c.close()
