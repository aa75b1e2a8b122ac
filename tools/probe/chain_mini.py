import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi, tower_layout as tl
L = _cabi.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
K, n, N = int(sys.argv[1]), 11, int(sys.argv[2])
mode = sys.argv[3]
ws = [(torch.randn(64, 64, 3, 3, device='cuda') * 0.05).to(torch.bfloat16) for _ in range(2 * K)]
wall = torch.cat([tl.pack_conv_weights(w) for w in ws]).contiguous()
ball = torch.zeros(2 * K * 64, device='cuda')
x = (torch.randn(N, n, n, 64, device='cuda') * 0.5).to(torch.bfloat16)
xa = tl.to_slabs(x)
torch.cuda.synchronize()
if mode == 'single':
    rc = L.az_nn_resblock(P(xa), P(wall), P(ball), None, n, N, st)
else:
    rc = L.az_nn_resblocks(P(xa), P(wall), P(ball), None, n, N, K, st)
print('rc', rc, flush=True)
torch.cuda.synchronize()
print('ok', mode, K, N)
