"""Global iterations of BASELINE configs[2] (256^3, subboxsize 10) restricted to its first `nsrc` sources -- with the
default 50 these are exactly the BB+QPL sources, which trace the whole box and carry 99 % of the workload's updates --
the command the round-2 ncu captures of the multi-SED sweep kernel are taken from.
usage: profile_cfg2.py [nsrc] [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import c2ray_b200
nsrc = int(sys.argv[1]) if len(sys.argv) > 1 else 50
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
p = c2ray_b200.synth.make_problem(3, n=256)
p["srcpos"] = np.ascontiguousarray(p["srcpos"][:nsrc]); p["NormFlux"] = np.ascontiguousarray(p["NormFlux"][:nsrc])
p["NormFluxQPL"] = np.ascontiguousarray(p["NormFluxQPL"][:nsrc])
c = c2ray_b200.from_problem(p, device=0)
c.begin_step()
for it in range(iters):
    c.set_rates_to_zero()
    l0 = c.sweep_launch_count()
    c.timer_start(); upd = c.pass_all_sources(it + 1, p["dt"]); ms_s = c.timer_stop()
    nl = c.sweep_launch_count() - l0
    c.timer_start(); cf = c.global_pass(p["dt"]); ms_c = c.timer_stop()
    print(f"iter {it+1}: {upd} updates in {ms_s:.2f} ms ({upd/ms_s/1e3:.1f} M/s), {nl} sweep launches; global pass {ms_c:.2f} ms "
          f"({c.N3/ms_c/1e3:.1f} Mcells/s) conv_flag={cf}")
c.close()
