"""SHA-256 of the rate grids and fractions after two global iterations in deterministic mode (bitwise reproducible):
two builds of the library that claim bit-identical arithmetic must print the same line.
usage: C2RAY_B200_LIB=lib_x.so python tools/rates_digest.py [mesh] [config]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import c2ray_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p = c2ray_b200.synth.make_problem(cfg, n=n, num_src=6)
c = c2ray_b200.from_problem(p, device=0, deterministic=True)
c.begin_step()
h = hashlib.sha256()
for it in range(2):
    c.set_rates_to_zero()
    upd = c.pass_all_sources(it + 1, p["dt"])
    for a in c.get_rates():
        h.update(np.ascontiguousarray(a).tobytes())
    c.global_pass(p["dt"])
    for a in c.get_work_state():
        h.update(np.ascontiguousarray(a).tobytes())
print(f"mesh {n} config {cfg}: {upd} updates, sha256 {h.hexdigest()}")
c.close()
