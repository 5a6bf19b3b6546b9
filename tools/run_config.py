"""Run global iterations of a BASELINE config on the GPU and print per-phase timings.
usage: run_config.py <config 1..4> [mesh] [num_src] [iterations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import c2ray_b200
cfg = int(sys.argv[1]); mesh = int(sys.argv[2]) if len(sys.argv) > 2 else None
nsrc = int(sys.argv[3]) if len(sys.argv) > 3 else None
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
t0 = time.time()
p = c2ray_b200.synth.make_problem(cfg, n=mesh, num_src=nsrc)
print(f"config {cfg}: mesh {p['mesh'][0]}^3, {len(p['NormFlux'])} sources, qpl sources {0 if p['NormFluxQPL'] is None else int((p['NormFluxQPL']>0).sum())}, setup {time.time()-t0:.1f}s")
c = c2ray_b200.from_problem(p, device=0)
c.begin_step()
for it in range(iters):
    c.set_rates_to_zero()
    c.timer_start(); upd = c.pass_all_sources(it + 1, p["dt"]); ms_s = c.timer_stop()
    c.timer_start(); cf = c.global_pass(p["dt"]); ms_c = c.timer_stop()
    r = c.get_rates()
    print(f"iter {it+1}: {upd} updates ({upd/len(p['NormFlux'])/c.N3:.4f} of full coverage) in {ms_s:.1f} ms ({upd/ms_s/1e3:.1f} M/s); "
          f"global pass {ms_c:.2f} ms ({c.N3/ms_c/1e3:.1f} Mcells/s) conv_flag={cf} max phih={r[0].max():.3e} finite={all(np.isfinite(a).all() for a in r)}")
xh_av = c.get_work_state()[0]
print("mean xHII_av", xh_av[1].mean())
c.close()
