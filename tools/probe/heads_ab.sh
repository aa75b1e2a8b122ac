#!/bin/bash
# A/B on one box: head convolutions fused into the tower's last epilogue (1) or as their own kernel (0)
for rep in 1 2; do for fh in 0 1; do
AZALEA_B200_FUSE_HEADS=$fh timeout 300 python bench.py --no-cpu-baseline --skip-configs --skip-tree-only --steps 6 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fuse_heads $fh rep $rep:', round(d['value']), 'sims/s', round(d['ms_per_step'],2), 'ms/step; e2e', round(d['e2e']['value']), '; tower', round(d['roofline']['avg_launch_ms'],3), 'ms; sm', d['clocks']['sm_mhz'])"
done; done
