#!/bin/bash
# A/B of builds of the library on one box: tools/probe/lib_ab.sh name=path.so [name=path.so ...]
# (alternating runs of the self-play step, python bench.py --steps 6)
for rep in 1 2; do for kv in "$@"; do
name=${kv%%=*}; lib=${kv#*=}
AZALEA_B200_LIB=$PWD/$lib timeout 200 python bench.py --no-cpu-baseline --skip-configs --skip-tree-only --steps 6 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$name rep $rep:', round(d['value']), 'sims/s', round(d['ms_per_step'],2), 'ms/step; e2e', round(d['e2e']['value']), '; tower', round(d['roofline']['avg_launch_ms'],3), 'ms frac', round(d['roofline']['frac'],3), '; sm', d['clocks']['sm_mhz'])"
done; done
