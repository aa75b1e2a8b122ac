import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi, tower_layout as tl
L = _cabi.lib()
n, N = 11, 1029
torch.manual_seed(111)
x = (torch.randn(N, n, n, 64, device='cuda') * 0.5).to(torch.bfloat16)
r = (torch.randn(N, n, n, 64, device='cuda') * 0.5).to(torch.bfloat16)
w = (torch.randn(64, 64, 3, 3, device='cuda') * 0.05).to(torch.bfloat16)
b = torch.randn(64, device='cuda') * 0.1
hw = torch.randn(6, 64, device='cuda') * 0.2
hb = torch.randn(6, device='cuda') * 0.1
xp, rp, wp = tl.to_slabs(x), tl.to_slabs(r), tl.pack_conv_weights(w)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: ctypes.c_void_p(t.data_ptr())
nn = n * n
stride = (nn * 6 + 7) // 8 * 8
out = torch.zeros_like(xp)
L.az_nn_conv3x3(p(xp), p(wp), p(b), p(rp), p(out), n, N, stream)
want = torch.zeros(N, stride, dtype=torch.bfloat16, device='cuda')
L.az_nn_heads(p(out), N * nn, p(hw), p(hb), p(want), stride, 64, 6, n, stream)
for rep in range(3):
    got = torch.zeros(N, stride, dtype=torch.bfloat16, device='cuda')
    L.az_nn_conv3x3_heads(p(xp), p(wp), p(b), p(rp), p(hw), p(hb), p(got), stride, 6, n, N, stream)
    torch.cuda.synchronize()
    d = (got[:, :nn * 6].float() - want[:, :nn * 6].float()).abs().view(N, n, n, 6)
    bad = (d > 0.02).nonzero()
    print('rep', rep, 'max', float(d.max()), 'bad entries', len(bad), 'boards', sorted(set(bad[:, 0].tolist()))[:12],
          'rows y', sorted(set(bad[:, 1].tolist())), 'x', sorted(set(bad[:, 2].tolist())), 'heads', sorted(set(bad[:, 3].tolist())))
