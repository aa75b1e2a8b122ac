import ctypes, os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi, tower_layout as tl
L = _cabi.lib()
torch.manual_seed(0)
dev = 'cuda'
for n, N in ((11, 8), (11, 4096), (19, 5), (11, 40960)):
    x = (torch.randn(N, n, n, 64, device=dev) * 0.5).to(torch.bfloat16)
    r = (torch.randn(N, n, n, 64, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(64, 64, 3, 3, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(64, device=dev) * 0.1
    xp, rp, wp = tl.to_slabs(x), tl.to_slabs(r), tl.pack_conv_weights(w)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    for use_res in (False, True):
        out = torch.zeros_like(xp)
        rc = L.az_nn_conv3x3(P(xp), P(wp), P(b), P(rp) if use_res else None, P(out), n, N, stream)
        torch.cuda.synchronize()
        got, rest = tl.from_slabs(out, n, N)
        want = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), b, padding=1)
        if use_res:
            want = want + r.permute(0, 3, 1, 2).float()
        want = F.relu(want).permute(0, 2, 3, 1)
        err = (got.float() - want).abs().max().item()
        print(f'n={n} N={N} resid={use_res} rc={rc} max_err={err:.4f} scale={want.abs().max().item():.2f} pad_max={rest}', flush=True)
    if N >= 4096:
        for use_res in (False, True):
            rr = P(rp) if use_res else None
            for _ in range(3):
                L.az_nn_conv3x3(P(xp), P(wp), P(b), rr, P(out), n, N, stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                L.az_nn_conv3x3(P(xp), P(wp), P(b), rr, P(out), n, N, stream)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f'   time resid={use_res}: {ms:.4f} ms  useful {2*n*n*64*576*N/ms/1e9:.1f} TFLOP/s', flush=True)
