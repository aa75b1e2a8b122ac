"""Probe: when do the CTAs of one chained k_resblock launch start and finish?

k_resblock gives each resident cluster a fixed, contiguous share of the board groups.  With
debug flag 16 every CTA writes %globaltimer at its start and end (and its %smid) into the probe
buffer; this prints the spread of the finishing times -- the part of the launch during which
some SMs already idle -- for the sizes of the self-play step.

    python tools/probe/cluster_times.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi, tower_layout as tl  # noqa: E402

L = _cabi.lib()
L.azb_set_prof.argtypes = [ctypes.c_void_p]
L.azb_set_prof.restype = None
L.azb_set_debug.argtypes = [ctypes.c_int]
L.azb_set_debug.restype = None
P = lambda t: ctypes.c_void_p(t.data_ptr())
torch.manual_seed(0)
SCR = torch.zeros(1 << 24, dtype=torch.uint8, device='cuda')
prof = torch.zeros(128 + 4 * 160, dtype=torch.int64, device='cuda')
blocks = 6
for n, N in ((11, 20480), (11, 40960), (19, 5120)):
    rows = L.az_nn_tower_rows(n, N)
    x = torch.zeros(rows, 64, device='cuda', dtype=torch.bfloat16)
    x[8:rows - 16] = (torch.rand(rows - 24, 64, device='cuda') * 0.1).to(torch.bfloat16)
    w = torch.cat([tl.pack_conv_weights((torch.randn(64, 64, 3, 3, device='cuda') * 0.02).to(torch.bfloat16))
                   for _ in range(2 * blocks)]).contiguous()
    b = torch.zeros(2 * blocks * 64, device='cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        rc = L.az_nn_resblocks(P(x), P(w), P(b), P(SCR), n, N, blocks, st)
        assert rc == 0, rc

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    for rep in range(3):
        prof.zero_()
        L.azb_set_prof(P(prof))
        L.azb_set_debug(16)
        run()
        torch.cuda.synchronize()
        L.azb_set_debug(0)
        L.azb_set_prof(None)
        t = prof[128:].view(-1, 4).cpu()
        t = t[t[:, 0] > 0]
        start, end, smid, nt = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
        t0 = int(start.min())
        dur = (end - t0).double() / 1e3        # us since the first CTA started
        sta = (start - t0).double() / 1e3
        d = dur.sort().values
        k = len(d)
        print(f'n={n} N={N} rep {rep}: {ms:.3f} ms/launch (events); {k} CTAs; start spread {float(sta.max()):.1f} us; '
              f'end: min {float(d[0]):.0f} p10 {float(d[k // 10]):.0f} median {float(d[k // 2]):.0f} '
              f'p90 {float(d[9 * k // 10]):.0f} max {float(d[-1]):.0f} us; mean/max {float(d.mean() / d[-1]):.3f}',
              flush=True)
        if rep == 2:
            # work and pace per cluster: slabs (NT) and us per slab
            pace = (end - start).double() / 1e3 / nt.double().clamp(min=1)
            order = dur.argsort()
            print('   slowest:', [(int(smid[i]), int(nt[i]), round(float(dur[i])), round(float(pace[i]), 4)) for i in order[-6:]])
            print('   fastest:', [(int(smid[i]), int(nt[i]), round(float(dur[i])), round(float(pace[i]), 4)) for i in order[:6]])
            print(f'   pace us/slab: min {float(pace.min()):.4f} median {float(pace.median()):.4f} max {float(pace.max()):.4f}; '
                  f'slabs per CTA: min {int(nt.min())} max {int(nt.max())}')
