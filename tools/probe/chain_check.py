"""az_nn_resblocks (K residual blocks chained in one launch) against K az_nn_resblock launches:
bit-identity at several sizes, repeated, and burst time at 20480 / 40960 boards."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi, tower_layout as tl
L = _cabi.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(3)
K = 6
for n, sizes in ((11, (1, 10, 64, 750, 4100, 4100, 20000)), (19, (6, 500)), (2, (42, 3000)), (5, (21, 1000))):
    ws = [(torch.randn(64, 64, 3, 3, device='cuda') * 0.05).to(torch.bfloat16) for _ in range(2 * K)]
    bs = [torch.randn(64, device='cuda') * 0.1 for _ in range(2 * K)]
    wall = torch.cat([tl.pack_conv_weights(w) for w in ws]).contiguous()
    ball = torch.cat(bs).contiguous()
    per_block = wall.numel() // K
    for N in sizes:
        x = (torch.randn(N, n, n, 64, device='cuda') * 0.5).to(torch.bfloat16)
        xa, xb = tl.to_slabs(x), tl.to_slabs(x)
        for b in range(K):
            rc = L.az_nn_resblock(P(xa), ctypes.c_void_p(wall.data_ptr() + b * per_block * wall.element_size()),
                                  ctypes.c_void_p(ball.data_ptr() + b * 128 * 4), None, n, N, st)
            assert rc == 0
        rc = L.az_nn_resblocks(P(xb), P(wall), P(ball), None, n, N, K, st)
        assert rc == 0, (rc, L.az_last_cuda_error())
        torch.cuda.synchronize()
        bad = int((xa.view(torch.int16) != xb.view(torch.int16)).any(1).sum())
        print(f'n={n} N={N}: rows differing {bad} of {xa.shape[0]}', flush=True)
n = 11
for N in (20480, 40960):
    rows = L.az_nn_tower_rows(n, N)
    x = torch.zeros(rows, 64, device='cuda', dtype=torch.bfloat16)
    x[8:rows - 16] = (torch.rand(rows - 24, 64, device='cuda') * 0.1).to(torch.bfloat16)
    ws = [(torch.randn(64, 64, 3, 3, device='cuda') * 0.02).to(torch.bfloat16) for _ in range(2 * K)]
    wall = torch.cat([tl.pack_conv_weights(w) for w in ws]).contiguous()
    ball = torch.zeros(2 * K * 64, device='cuda')
    per_block = wall.numel() // K

    def six():
        for b in range(K):
            L.az_nn_resblock(P(x), ctypes.c_void_p(wall.data_ptr() + b * per_block * wall.element_size()),
                             ctypes.c_void_p(ball.data_ptr() + b * 128 * 4), None, n, N, st)

    def chain():
        L.az_nn_resblocks(P(x), P(wall), P(ball), None, n, N, K, st)

    # A/B interleaved (the clocks sag under sustained load: whoever runs second looks slower)
    res = {'6 launches': [], 'chained': []}
    for fn in (six, chain):
        for _ in range(10):
            fn()
    for rnd in range(8):
        for name, fn in (('6 launches', six), ('chained', chain)):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / 10)
    for name, v in res.items():
        v = sorted(v)
        print(f'N={N} {name}: median {v[len(v) // 2]:.4f} ms per tower (min {v[0]:.4f}, max {v[-1]:.4f}; {v[len(v) // 2] / K:.4f} per block)', flush=True)
