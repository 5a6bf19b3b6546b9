#!/bin/bash
# A/B timing of tuning builds of libc2ray_b200.so (lib_<tag>.so next to the product library; C2RAY_B200_LIB selects one).
# usage: tools/ab_variants.sh tag1 tag2 ...   -> one line per build: RT pass of configs[1] (128^3, 16 sources, iterations 3-4),
# configs[0] full step, configs[2] iterations 2-3 (256^3, 1000 sources)
cd "$(dirname "$0")/.."
for t in "$@"; do
  lib=lib_$t.so; [ "$t" = main ] && lib=libc2ray_b200.so
  a=$(C2RAY_B200_LIB=$lib python tools/profile_step.py 128 4 2>&1 | tail -2 | sed -e 's/ updates in / /' | awk '{printf "%s ms ", $4}')
  b=$(C2RAY_B200_LIB=$lib python tools/bench_configs.py --only 0,2 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: continue
    if d['config'] == 0: print('cfg0 step %.1f ms sweeps %.1f ms' % (d['s_per_timestep'] * 1e3, d['ms_sweep']), end=' ')
    if d['config'] == 2: print('cfg2 pass %.1f ms %.3f G/s' % (d['ms_sweep'], d['sweep_updates_per_s'] / 1e9), end=' ')
")
  echo "$t: cfg1 passes $a| $b"
done
