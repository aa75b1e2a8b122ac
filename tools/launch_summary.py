"""Aggregate an ncu launch list (gpu__time_duration.sum per launch) by kernel."""
import collections
import csv
import sys

src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i + 1
        break
kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(',', ''))
    except ValueError:
        continue
    name = r[kn].split('(')[0][:84]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if 'k_' in k and ('az' in k or k.strip().startswith(('k_', 'void k_'))))
lines = [f'# {title}',
         '# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and',
         '# serialised: compare SHARES, not absolutes',
         f'# window: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} us total; our kernels '
         f'{100 * ours / tot:.1f}% of it', '']
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f'{name:84s} n {n:4d}  total {t / 1e3:10.1f} us  avg {t / n / 1e3:8.1f} us  share {100 * t / tot:5.1f}%')
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
