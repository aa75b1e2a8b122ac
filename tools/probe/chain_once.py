"""A few launches of az_nn_resblocks (6 blocks chained) on 40960 boards 11x11 (for ncu --set full -k regex:k_resblock)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi, tower_layout as tl
L = _cabi.lib()
n, N, K = 11, int(sys.argv[1]) if len(sys.argv) > 1 else 40960, 6
rows = L.az_nn_tower_rows(n, N)
x = torch.zeros(rows, 64, device='cuda', dtype=torch.bfloat16)
x[8:rows - 16] = (torch.rand(rows - 24, 64, device='cuda') * 0.1).to(torch.bfloat16)
w = torch.cat([tl.pack_conv_weights((torch.randn(64, 64, 3, 3, device='cuda') * 0.02).to(torch.bfloat16)) for _ in range(2 * K)]).contiguous()
b = torch.zeros(128 * K, device='cuda')
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    L.az_nn_resblocks(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr()),
                      ctypes.cast(b.data_ptr(), ctypes.POINTER(ctypes.c_float)), None, n, N, K, s)
torch.cuda.synchronize()
