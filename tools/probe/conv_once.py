import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi
L = _cabi.lib()
n, N = 11, 40960
rows = L.az_nn_tower_rows(n, N)
x = (torch.randn(rows, 64, device='cuda') * 0.5).to(torch.bfloat16)
r = (torch.randn(rows, 64, device='cuda') * 0.5).to(torch.bfloat16)
w = (torch.randn(9 * 64, 64, device='cuda') * 0.05).to(torch.bfloat16)
b = torch.randn(64, device='cuda') * 0.1
out = torch.zeros_like(x)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(use_res):
    L.az_nn_conv3x3(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(w.data_ptr()), ctypes.c_void_p(b.data_ptr()),
                    ctypes.c_void_p(r.data_ptr()) if use_res else None, ctypes.c_void_p(out.data_ptr()), n, N, s)
for use_res in (0, 1, 0, 1):
    run(use_res)
torch.cuda.synchronize()
if os.environ.get('AZT_TIME'):
    for use_res in (0, 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run(use_res)
        e1.record(); torch.cuda.synchronize()
        print('debug', os.environ.get('AZT_DEBUG', '0'), 'resid', use_res, 'ms', e0.elapsed_time(e1) / 20, flush=True)
