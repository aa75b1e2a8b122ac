"""Two streams, each running a stream of k_conv3x3 launches on its own buffers: do concurrent
launches of the kernel interfere?"""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi
L = _cabi.lib()
n, N = 11, 20480
which = sys.argv[1] if len(sys.argv) > 1 else 'conv'
rows = L.az_nn_tower_rows(n, N)
streams = [torch.cuda.Stream() for _ in range(2)]
bufs = []
for s in streams:
    x = torch.zeros(rows, 64, device='cuda', dtype=torch.bfloat16)
    y = torch.zeros_like(x)
    z = torch.zeros_like(x)
    w = (torch.randn(9 * 64, 64, device='cuda') * 0.02).to(torch.bfloat16)
    b = torch.zeros(64, device='cuda')
    bufs.append((x, y, w, b, z))
torch.cuda.synchronize()
P = lambda t: ctypes.c_void_p(t.data_ptr())
for it in range(40):
    for s, (x, y, w, b, z) in zip(streams, bufs):
        with torch.cuda.stream(s):
            st = ctypes.c_void_p(s.cuda_stream)
            for _ in range(6):
                if which in ('conv', 'plain'):
                    L.az_nn_conv3x3(P(x), P(w), P(b), None, P(y), n, N, st)
                if which in ('conv', 'resid'):
                    L.az_nn_conv3x3(P(y), P(w), P(b), P(x), P(x), n, N, st)
                if which == 'resid_oop':
                    L.az_nn_conv3x3(P(y), P(w), P(b), P(x), P(z), n, N, st)
    torch.cuda.synchronize()
    if it % 10 == 9: print(which, 'iteration', it + 1, 'ok', flush=True)
print('done', flush=True)
