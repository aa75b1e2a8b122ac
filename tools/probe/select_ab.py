"""A/B of k_select builds on one box: alternate tree-only bench runs (stub evaluator, steady
state, CUDA graph) of several builds of the library (AZALEA_B200_LIB), and count k_select's warp
instructions per descent under ncu at one fixed launch.

    python tools/probe/select_ab.py name=path.so [name=path.so ...] [--reps 2] [--ncu]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
args = [a for a in sys.argv[1:] if '=' in a]
reps = int(sys.argv[sys.argv.index('--reps') + 1]) if '--reps' in sys.argv else 2
libs = [a.split('=', 1) for a in args]
for rep in range(reps):
    for name, path in libs:
        env = dict(os.environ, AZALEA_B200_LIB=os.path.join(ROOT, path))
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--evaluator', 'stub',
                              '--no-cpu-baseline', '--skip-configs', '--steps', '8'],
                             env=env, capture_output=True, text=True).stdout
        d = json.loads(out.strip().splitlines()[-1])
        print(f'{name:12s} rep {rep}: tree-only {d["value"]:.4e} sims/s, {d["ms_per_step"]:.2f} ms/step, '
              f'k_select {d["roofline"]["avg_launch_ms"] * 1e3:.1f} us/launch, depth {d["mean_depth"]:.2f}', flush=True)
if '--ncu' in sys.argv:
    for name, path in libs:
        env = dict(os.environ, AZALEA_B200_LIB=os.path.join(ROOT, path))
        out = subprocess.run(['ncu', '--metrics', 'smsp__inst_executed.sum,gpu__time_duration.sum',
                              '-k', 'regex:k_select', '--launch-skip', '120', '-c', '2', sys.executable,
                              os.path.join(ROOT, 'tools', 'profile_step.py'), '--evaluator', 'stub', '--warm', '1'],
                             env=env, capture_output=True, text=True).stdout
        vals = [l.split()[-1] for l in out.splitlines() if 'inst_executed' in l or 'duration' in l]
        print(f'{name:12s} ncu (launch 120, 121): duration us / warp inst: {vals}; per descent '
              f'{[round(float(v.replace(",", "")) / 40960) for v in vals[1::2]]}', flush=True)
