"""Probe: where the time of a two-stream lockstep move goes.

Runs one eager move (streams=2) with a CUDA event recorded after every kernel
launch on both streams (a proxy around the ctypes library + torch.mm), then
prints the merged end-time line of a few search batches and a summary: how
long the residual towers of the two windows keep the device busy and how much
of the move no tower is running (the exposed part of the small kernels).

An event gives the time its stream REACHED it, i.e. the end of the kernel
before it; a tower is taken to start when its stem has ended and the other
window's previous tower has ended (towers fill the device, so they serialise).

    python tools/probe/timeline.py [--games 4096] [--streams 2] [--out gpurun_out/timeline.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from azalea_b200 import LockstepSelfPlay, _cabi  # noqa: E402
from azalea_b200.network import HexNetwork  # noqa: E402

WATCH = {'az_mcts_select': 'select', 'az_mcts_expand_backup': 'expand', 'az_nn_stem': 'stem',
         'az_nn_resblocks': 'tower', 'az_nn_heads': 'heads', 'az_nn_tail': 'tail',
         'az_mcts_select_root': 'select_root', 'az_mcts_expand_root': 'expand_root',
         'az_play_commit': 'commit'}


class Recorder:
    def __init__(self, real):
        self._real = real
        self.on = False
        self.log = []       # (stream id, name, event)

    def mark(self, name):
        if self.on:
            s = torch.cuda.current_stream()
            e = torch.cuda.Event(enable_timing=True)
            e.record(s)
            self.log.append((s.cuda_stream, name, e))

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        tag = WATCH.get(name)
        if tag is None:
            return fn

        def wrapped(*a):
            if tag == 'tower':
                self.mark('tower_ready')
            rc = fn(*a)
            self.mark(tag)
            return rc
        return wrapped


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--games', type=int, default=4096)
    ap.add_argument('--streams', type=int, default=2)
    ap.add_argument('--board', type=int, default=11)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'timeline.json'))
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    net = HexNetwork(args.board, 6, 64).eval().to(dev)
    net.prepare_inference(torch.bfloat16)
    rec = Recorder(_cabi.lib())
    _cabi._lib = rec
    real_mm = torch.mm

    def mm(*a, **k):
        r = real_mm(*a, **k)
        rec.mark('gemm')
        return r
    torch.mm = mm
    sp = LockstepSelfPlay(net, num_games=args.games, board_size=args.board, simulations=800,
                          search_batch_size=10, cuda_graph=False, streams=args.streams, device=dev,
                          seed=1)
    for _ in range(3):
        sp.step_move()
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True)
    base.record(torch.cuda.current_stream())
    rec.on = True
    sp.step_move()
    rec.on = False
    torch.cuda.synchronize()
    rows = [(sid, name, base.elapsed_time(e)) for sid, name, e in rec.log]
    sids = sorted({r[0] for r in rows})
    rows = [(sids.index(s), n, t) for s, n, t in rows]
    total = max(t for _, _, t in rows)
    # towers in end-time order; start = max(own stem end, previous tower end)
    towers = []
    ready = {}
    for s, n, t in rows:
        if n == 'tower_ready':
            ready[s] = t
        elif n == 'tower':
            towers.append((t, s, ready[s]))
    towers.sort()
    busy, prev_end, durs, waits = 0.0, 0.0, [], []
    for end, s, rdy in towers:
        start = max(rdy, prev_end)
        durs.append(end - start)
        waits.append(max(0.0, prev_end - rdy))
        busy += end - start
        prev_end = end
    merged = sorted(rows, key=lambda r: r[2])
    mid = len(merged) // 2
    print('move %.2f ms, %d towers, tower time %.2f ms (%.1f %%), no tower running %.2f ms' %
          (total, len(towers), busy, 100 * busy / total, total - busy))
    durs_s = sorted(durs)
    print('tower duration (start = stem end or previous tower end): median %.3f, p10 %.3f, p90 %.3f ms'
          % (durs_s[len(durs) // 2], durs_s[len(durs) // 10], durs_s[9 * len(durs) // 10]))
    print('tower waited for the other tower: mean %.3f ms' % (sum(waits) / len(waits)))
    # per-kernel time on its own stream = end - previous end on that stream (includes waiting for SMs)
    last = {}
    per = {}
    for s, n, t in rows:
        if s in last and n != 'tower_ready':
            per.setdefault(n, []).append(t - last[s])
        last[s] = t
    for n, v in sorted(per.items()):
        v = sorted(v)
        print('  %-12s n %4d  median %.3f  mean %.3f ms (stream-local gap to the previous end)' %
              (n, len(v), v[len(v) // 2], sum(v) / len(v)))
    print('merged end times around the middle of the move:')
    for s, n, t in merged[mid:mid + 40]:
        print('  %9.3f ms  %s%s' % (t, '            ' * s, n))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({'rows': rows, 'total_ms': total, 'tower_ms': busy}, open(args.out, 'w'))


if __name__ == '__main__':
    main()
