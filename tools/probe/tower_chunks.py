"""Does the tower run out of L2 when the batch is processed in chunks?  Times 12 convolutions
(6 residual blocks) over 40960 boards, chunk by chunk, for several chunk sizes."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi
L = _cabi.lib()
n, N = 11, 40960
bpg = L.az_nn_tower_group(n)
rows = L.az_nn_tower_rows(n, N)
x = (torch.randn(rows, 64, device='cuda') * 0.5).to(torch.bfloat16)
y = torch.zeros_like(x)
ws = [(torch.randn(9 * 64, 64, device='cuda') * 0.02).to(torch.bfloat16) for _ in range(12)]
b = torch.zeros(64, device='cuda')
P = lambda t, off=0: ctypes.c_void_p(t.data_ptr() + off)
def tower(chunk):
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for c0 in range(0, N, chunk):
        nb = min(chunk, N - c0)
        off = (c0 // bpg) * n * 128 * 128
        for i in range(6):
            L.az_nn_conv3x3(P(x, off), P(ws[2 * i]), P(b), None, P(y, off), n, nb, s)
            L.az_nn_conv3x3(P(y, off), P(ws[2 * i + 1]), P(b), P(x, off), P(x, off), n, nb, s)
for chunk in (40960, 20480, 10240, 5120, 2960, 2560, 1480, 1280):
    g = torch.cuda.CUDAGraph()
    tower(chunk); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        tower(chunk)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f'chunk {chunk:6d} boards ({chunk // bpg:5d} groups, {chunk * 121 * 128 * 2 / 1e6:6.1f} MB x+y): tower {e0.elapsed_time(e1) / 5:.3f} ms', flush=True)
