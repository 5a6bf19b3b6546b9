#!/bin/bash
# A/B timing of run-time switches of libc2ray_b200.so: every argument is an environment assignment list (quoted), e.g.
#   tools/ab_env.sh "C2RAY_SWEEP_LANES_MODE=0" "C2RAY_SWEEP_LANES_MODE=1 C2RAY_SWEEP_LANES_FILL=2"
# one line per setting: configs[1] RT passes (iterations 3-4), configs[0] full step, configs[2] iterations 2-3
cd "$(dirname "$0")/.."
for e in "$@"; do
  a=$(env $e python tools/profile_step.py 128 4 2>&1 | tail -2 | sed -e 's/ updates in / /' | awk '{printf "%s ms ", $4}')
  b=$(env $e python tools/bench_configs.py --only ${AB_CONFIGS:-0,2} 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    if d['config'] == 0: print('cfg0 step %.1f ms sweeps %.1f ms' % (d['s_per_timestep'] * 1e3, d['ms_sweep']), end=' ')
    if d['config'] == 2: print('cfg2 pass %.1f ms %.3f G/s' % (d['ms_sweep'], d['sweep_updates_per_s'] / 1e9), end=' ')
    if d['config'] == 3: print('cfg3 pass %.2f ms %.3f G/s' % (d['ms_sweep'], d['sweep_updates_per_s'] / 1e9), end=' ')
")
  echo "[$e]: cfg1 passes $a| $b"
done
