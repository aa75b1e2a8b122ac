"""Where does az_nn_resblock differ from the two-launch path?  Probe mode: the intermediate y
is also written to global memory and the residual ring is checked against global memory."""
import collections, ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi, tower_layout as tl
L = _cabi.lib()
L.azb_set_debug.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
L.azb_set_debug.restype = None
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def report(tag, a, b):
    a, b = tl.unswizzle_rows(a).float(), tl.unswizzle_rows(b).float()
    diff = a != b
    rows = diff.any(1).nonzero().flatten()
    print(f'  {tag}: rows differing {len(rows)} of {a.shape[0]}, max abs diff {float((a - b).abs().max()):.3f}')
    if len(rows):
        q, l = (rows - 8) // 128, (rows - 8) % 128
        print('    slabs (q: rows):', sorted(collections.Counter(q.tolist()).items())[:24])
        hist = collections.Counter((l // 32).tolist())
        print('    rows by warp quadrant:', sorted(hist.items()))
        ch = diff[rows].reshape(len(rows), 8, 8).any(2).float().sum(0)
        print('    rows with a difference, per 16-byte chunk:', [int(v) for v in ch.tolist()])


for n, N, dbg in ((11, 64, False), (11, 64, True), (11, 4100, True), (11, 4100, False)):
    torch.manual_seed(1)
    x = (torch.randn(N, n, n, 64, device='cuda') * 0.5).to(torch.bfloat16)
    ws = [(torch.randn(64, 64, 3, 3, device='cuda') * 0.05).to(torch.bfloat16) for _ in range(2)]
    bs = [torch.randn(64, device='cuda') * 0.1 for _ in range(2)]
    wp = [tl.pack_conv_weights(w) for w in ws]
    w12, b12 = torch.cat(wp).contiguous(), torch.cat(bs).contiguous()
    xa, ya = tl.to_slabs(x), torch.zeros_like(tl.to_slabs(x))
    L.az_nn_conv3x3(P(xa), P(wp[0]), P(bs[0]), None, P(ya), n, N, st)
    L.az_nn_conv3x3(P(ya), P(wp[1]), P(bs[1]), P(xa), P(xa), n, N, st)
    torch.cuda.synchronize()
    print(f'n={n} N={N} probe={dbg}')
    for rep in range(3):
        xb = tl.to_slabs(x)
        yb = torch.zeros_like(xb)
        cnt = torch.zeros(4, dtype=torch.int32, device='cuda')
        torch.cuda.synchronize()
        L.azb_set_debug(P(yb) if dbg else None, P(cnt) if dbg else None)
        rc = L.az_nn_resblock(P(xb), P(w12), P(b12), n, N, st)
        torch.cuda.synchronize()
        L.azb_set_debug(None, None)
        print(f' rep {rep}: rc={rc} residual-ring mismatching chunks: {int(cnt[0])}')
        if dbg:
            report('intermediate y', ya, yb)
        report('output', xa, xb)
