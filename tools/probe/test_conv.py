import ctypes, os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi
L = _cabi.lib()
HALO = 16

def swz_rows(t):            # t [R, 64] -> chunk j of row R at j ^ (R & 7)
    R = t.shape[0]
    v = t.view(R, 8, 8)
    idx = (torch.arange(8, device=t.device)[None, :] ^ (torch.arange(R, device=t.device)[:, None] & 7))
    out = torch.empty_like(v)
    out.scatter_(1, idx[:, :, None].expand(R, 8, 8), v)
    return out.view(R, 64)

def unswz_rows(t):
    R = t.shape[0]
    v = t.view(R, 8, 8)
    idx = (torch.arange(8, device=t.device)[None, :] ^ (torch.arange(R, device=t.device)[:, None] & 7))
    return torch.gather(v, 1, idx[:, :, None].expand(R, 8, 8)).reshape(R, 64)

def to_padded(x):           # x [N, n, n, 64] bf16
    N, n = x.shape[0], x.shape[1]
    HALO = L.az_nn_tower_halo(n)
    p = torch.zeros(N, n + 1, n + 1, 64, dtype=x.dtype, device=x.device)
    p[:, :n, :n] = x
    buf = torch.zeros(L.az_nn_tower_rows(n, N), 64, dtype=x.dtype, device=x.device)
    buf[HALO:HALO + N * (n + 1) ** 2] = p.view(-1, 64)
    return swz_rows(buf)

def from_padded(buf, N, n):
    HALO = L.az_nn_tower_halo(n)
    t = unswz_rows(buf)[HALO:HALO + N * (n + 1) ** 2].view(N, n + 1, n + 1, 64)
    return t

def pack_w(w):              # [co, ci, 3, 3] -> [9, co, 64] swizzled by co
    t = w.permute(2, 3, 0, 1).reshape(9 * 64, 64).contiguous()
    return swz_rows(t).view(9, 64, 64)      # (tap*64 + co) & 7 == co & 7

torch.manual_seed(0)
dev = 'cuda'
for n, N in ((11, 8), (11, 4096), (19, 5), (11, 40960)):
    nb = L.az_nn_tower_group(n)
    N = (N + nb - 1) // nb * nb
    x = (torch.randn(N, n, n, 64, device=dev) * 0.5).to(torch.bfloat16)
    r = (torch.randn(N, n, n, 64, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(64, 64, 3, 3, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(64, device=dev) * 0.1
    xp, rp, wp = to_padded(x), to_padded(r), pack_w(w)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for use_res in (False, True):
        out = torch.full_like(xp, 7.0)
        H = L.az_nn_tower_halo(n)
        out[:H] = 0; out[H + N * (n + 1) ** 2:] = 0
        rc = L.az_nn_conv3x3(ctypes.c_void_p(xp.data_ptr()), ctypes.c_void_p(wp.data_ptr()),
                             ctypes.c_void_p(b.data_ptr()),
                             ctypes.c_void_p(rp.data_ptr()) if use_res else None,
                             ctypes.c_void_p(out.data_ptr()), n, N, stream)
        torch.cuda.synchronize()
        got = from_padded(out, N, n)
        want = F.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), b, padding=1)
        if use_res:
            want = want + r.permute(0, 3, 1, 2).float()
        want = F.relu(want).permute(0, 2, 3, 1)
        err = (got[:, :n, :n].float() - want).abs().max().item()
        pads = max(got[:, n].abs().max().item(), got[:, :, n].abs().max().item())
        print(f'n={n} N={N} nb={nb} resid={use_res} rc={rc} max_err={err:.4f} scale={want.abs().max().item():.2f} pad_max={pads}', flush=True)
    if N >= 4096:
        import time
        for use_res in (False, True):
            rr = ctypes.c_void_p(rp.data_ptr()) if use_res else None
            for _ in range(3):
                L.az_nn_conv3x3(ctypes.c_void_p(xp.data_ptr()), ctypes.c_void_p(wp.data_ptr()), ctypes.c_void_p(b.data_ptr()), rr, ctypes.c_void_p(out.data_ptr()), n, N, stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                L.az_nn_conv3x3(ctypes.c_void_p(xp.data_ptr()), ctypes.c_void_p(wp.data_ptr()), ctypes.c_void_p(b.data_ptr()), rr, ctypes.c_void_p(out.data_ptr()), n, N, stream)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f'   time resid={use_res}: {ms:.4f} ms  useful {2*n*n*64*576*N/ms/1e9:.1f} TFLOP/s', flush=True)
