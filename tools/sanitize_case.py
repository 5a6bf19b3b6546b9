"""Small end-to-end case for compute-sanitizer: a 12^3 three-source (BB + QPL) evolve3D step plus the batch hooks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import c2ray_b200
p = c2ray_b200.synth.make_problem(3, n=12, num_src=3)
c = c2ray_b200.from_problem(p, device=0)
st = c.evolve3D(0.0, p["dt"], 0)
print("niter", st["niter"], "updates", st["rt_updates"])
rng = np.random.default_rng(0)
col6 = 10.0 ** rng.uniform(12, 20, (64, 6)); col6[:, 1::2] += col6[:, 0::2]
c.photoion_rates(col6, np.full(64, 1e66), [1e5, 0, 1e3], np.full(64, 0.1))
c.ini_rec_colion_factors(np.array([1e3, 1e4, 1e5]))
c.close()
