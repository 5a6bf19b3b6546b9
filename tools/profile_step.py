"""One global iteration (RT pass over all sources + global chemistry pass) of BASELINE configs[1] -- the command the
ncu captures under profiles/ are taken from.  usage: profile_step.py [mesh] [iterations] [iso]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import c2ray_b200
mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
iso = len(sys.argv) > 3 and sys.argv[3] == "iso"
p = c2ray_b200.synth.make_problem(2, n=mesh, num_src=16, isothermal=iso)
c = c2ray_b200.from_problem(p, device=0)
c.begin_step()
for it in range(iters):
    c.set_rates_to_zero()
    c.timer_start(); upd = c.pass_all_sources(it + 1, p["dt"]); ms_s = c.timer_stop()
    c.timer_start(); cf = c.global_pass(p["dt"]); ms_c = c.timer_stop()
    print(f"iter {it+1}: {upd} updates in {ms_s:.2f} ms ({upd/ms_s/1e3:.1f} M/s); global pass {ms_c:.2f} ms ({c.N3/ms_c/1e3:.1f} Mcells/s) conv_flag={cf}")
c.close()
