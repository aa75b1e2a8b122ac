"""Summarise an .ncu-rep (read here, no GPU needed) into a short text file.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.txt [top_lines]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'lts__t_bytes.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
]


def ncu(rep, *args):
    return subprocess.run(['ncu', '-i', rep, *args], capture_output=True,
                          text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    lines = [f'# ncu summary of {rep} (ncu --set full --clock-control none)', '']
    rows = list(csv.reader(io.StringIO(ncu(rep, '--page', 'raw', '--csv'))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        lines.append('kernel: ' + r[hdr.index('Kernel Name')])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f'  {k:78s} {r[i]:>16s} {units[i]}')
        rd = float(r[hdr.index('dram__bytes_read.sum')])
        wr = float(r[hdr.index('dram__bytes_write.sum')])
        scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
        rd *= scale[units[hdr.index('dram__bytes_read.sum')]]
        wr *= scale[units[hdr.index('dram__bytes_write.sum')]]
        lines.append(f'  traffic (dram read + write) per launch: {rd + wr:.0f} bytes')
        lines.append('')
    src = list(csv.reader(io.StringIO(
        ncu(rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'))))
    cur, h, agg = None, None, {}
    for r in src:
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
        elif len(r) >= 2 and r[0] == 'Line No':
            h = r
            ii, si = h.index('Instructions Executed'), h.index('# Samples')
        elif h and len(r) > ii and r[0].isdigit():
            try:
                key = (cur, int(r[0]), r[1].strip()[:90])
                n, s = int(r[ii]), int(r[si])
            except ValueError:
                continue
            a = agg.setdefault(key, [0, 0])
            a[0] += n
            a[1] += s
    tot = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    lines.append(f'hottest source lines (share of warp instructions / of stall samples), '
                 f'total inst {tot}')
    for (f, ln, text), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        lines.append(f'  {100 * n / tot:5.1f}% inst {100 * s / ts:5.1f}% smp  {f}:{ln:<4d} {text}')
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
