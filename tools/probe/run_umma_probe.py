import ctypes, os, subprocess, sys, torch
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, 'umma_probe.so')
L = ctypes.CDLL(so)
torch.manual_seed(0)
a_rows = 256
A = (torch.randn(a_rows, 64, device='cuda')).to(torch.bfloat16)
B = (torch.randn(64, 64, device='cuda')).to(torch.bfloat16)
for r0 in (0, 8, 1, 13, 27):
    for ubo in (0, 1):
        D = torch.zeros(128, 64, device='cuda')
        rc = L.umma_probe(ctypes.c_void_p(A.data_ptr()), a_rows, r0, ctypes.c_void_p(B.data_ptr()),
                          ctypes.c_void_p(D.data_ptr()), ubo)
        want = A[r0:r0 + 128].float() @ B.float().t()
        err = (D - want).abs().max().item()
        print(f'r0={r0:3d} base_offset_used={ubo} rc={rc} max_err={err:.4f} ref_scale={want.abs().max().item():.2f}', flush=True)
