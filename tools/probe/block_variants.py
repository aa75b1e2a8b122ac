"""Correctness + time of k_resblock built with different ring depths (tools/probe/libvariant_*.so)."""
import ctypes, glob, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from azalea_b200 import _cabi, tower_layout as tl
L0 = _cabi.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())
SCR = torch.zeros(1 << 24, dtype=torch.uint8, device='cuda')     # >= az_nn_resblock_scratch_bytes()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
libs = [('default', L0)] + [(os.path.basename(p), ctypes.CDLL(p)) for p in sorted(glob.glob(os.path.join(ROOT, 'tools/probe/libvariant_*.so')))]
n = 11
torch.manual_seed(1)
for name, L in libs:
    L.az_nn_resblock.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p]
    bad = []
    for N in (64, 4100, 4100, 4100, 20000):
        x = (torch.randn(N, n, n, 64, device='cuda') * 0.5).to(torch.bfloat16)
        ws = [(torch.randn(64, 64, 3, 3, device='cuda') * 0.05).to(torch.bfloat16) for _ in range(2)]
        bs = [torch.randn(64, device='cuda') * 0.1 for _ in range(2)]
        wp = [tl.pack_conv_weights(w) for w in ws]
        w12, b12 = torch.cat(wp).contiguous(), torch.cat(bs).contiguous()
        xa, ya = tl.to_slabs(x), torch.zeros_like(tl.to_slabs(x))
        L0.az_nn_conv3x3(P(xa), P(wp[0]), P(bs[0]), None, P(ya), n, N, st)
        L0.az_nn_conv3x3(P(ya), P(wp[1]), P(bs[1]), P(xa), P(xa), n, N, st)
        xb = tl.to_slabs(x)
        rc = L.az_nn_resblock(P(xb), P(w12), P(b12), P(SCR), n, N, st)
        torch.cuda.synchronize()
        bad.append(int((xa.view(torch.int16) != xb.view(torch.int16)).any(1).sum()))
    N = 40960
    rows = L0.az_nn_tower_rows(n, N)
    x = torch.zeros(rows, 64, device='cuda', dtype=torch.bfloat16)
    x[8:rows - 16] = (torch.rand(rows - 24, 64, device='cuda') * 0.1).to(torch.bfloat16)
    for _ in range(3):
        L.az_nn_resblock(P(x), P(w12), P(b12), P(SCR), n, N, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        L.az_nn_resblock(P(x), P(w12), P(b12), P(SCR), n, N, st)
    e1.record()
    torch.cuda.synchronize()
    print(f'{name}: rc={rc} rows differing per run {bad}; {e0.elapsed_time(e1) / 30:.4f} ms per block (40960 boards)', flush=True)
