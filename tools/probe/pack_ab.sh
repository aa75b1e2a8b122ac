#!/bin/bash
# A/B on one box: the self-play step with and without packed leaves (bench.py --pack-leaves)
for rep in 1 2; do for pk in 0 1; do
timeout 300 python bench.py --no-cpu-baseline --skip-configs --skip-tree-only --steps 4 --pack-leaves $pk 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pack $pk rep $rep:', round(d['value']), 'sims/s', round(d['ms_per_step'],2), 'ms/step; tower', round(d['roofline']['avg_launch_ms'],3), 'ms; select', round(d['roofline_tree']['avg_launch_ms']*1e3,1), 'us; expand', round(d['roofline_tree']['expand_backup_avg_launch_ms']*1e3,1), 'us; sm', d['clocks']['sm_mhz'])"
done; done
