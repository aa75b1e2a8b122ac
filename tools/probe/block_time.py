"""Time one residual block of the tower: the fused launch (az_nn_resblock, csrc/az_block.cuh)
against the two az_nn_conv3x3 launches, burst clocks, CUDA events, 40960 boards 11x11 (and 19x19)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi, tower_layout as tl
L = _cabi.lib()
torch.manual_seed(0)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for n, N in ((11, 40960), (19, 5120), (11, 2048)):
    rows = L.az_nn_tower_rows(n, N)
    x = torch.zeros(rows, 64, device='cuda', dtype=torch.bfloat16)
    x[8:rows - 16] = (torch.rand(rows - 24, 64, device='cuda') * 0.1).to(torch.bfloat16)
    y = torch.zeros_like(x)
    w = [tl.pack_conv_weights((torch.randn(64, 64, 3, 3, device='cuda') * 0.02).to(torch.bfloat16)) for _ in range(2)]
    b = [torch.zeros(64, device='cuda') for _ in range(2)]
    w12, b12 = torch.cat(w).contiguous(), torch.cat(b).contiguous()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def two():
        L.az_nn_conv3x3(P(x), P(w[0]), P(b[0]), None, P(y), n, N, st)
        L.az_nn_conv3x3(P(y), P(w[1]), P(b[1]), P(x), P(x), n, N, st)

    def one():
        rc = L.az_nn_resblock(P(x), P(w12), P(b12), n, N, st)
        assert rc == 0, (rc, L.az_last_cuda_error())

    for name, fn in (('two launches', two), ('fused', one)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        reps = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flop = 2 * 2 * n * n * 64 * 576 * N
        print(f'n={n} N={N} {name}: {ms:.4f} ms per block, {flop / ms / 1e9:.0f} useful TFLOP/s', flush=True)
print('resident clusters:', L.az_nn_resblock_clusters())
