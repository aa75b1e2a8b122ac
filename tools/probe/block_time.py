"""Time one residual block of the tower: the fused launch (az_nn_resblock, csrc/az_block.cuh)
against the two az_nn_conv3x3 launches, burst clocks, CUDA events, 40960 boards 11x11 (and 19x19)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from azalea_b200 import _cabi, tower_layout as tl
L = _cabi.lib()
torch.manual_seed(0)
P = lambda t: ctypes.c_void_p(t.data_ptr())
SCR = torch.zeros(1 << 24, dtype=torch.uint8, device='cuda')     # >= az_nn_resblock_scratch_bytes()
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
        rc = L.az_nn_resblock(P(x), P(w12), P(b12), P(SCR), n, N, st)
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
# what costs what: parts of the fused kernel switched off (results are garbage, times are not)
n, N = 11, 40960
rows = L.az_nn_tower_rows(n, N)
x = torch.zeros(rows, 64, device='cuda', dtype=torch.bfloat16)
x[8:rows - 16] = (torch.rand(rows - 24, 64, device='cuda') * 0.1).to(torch.bfloat16)
for flags, what in ((0, 'everything'), (2, 'no global stores (C)'), (4, 'no residual loads (C)'),
                    (6, 'no global stores, no residual loads'), (8, 'no MMAs'), (14, 'barriers + TMEM + staging + copies only')):
    L.azb_set_debug(flags)
    for _ in range(3):
        L.az_nn_resblock(P(x), P(w12), P(b12), P(SCR), n, N, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        L.az_nn_resblock(P(x), P(w12), P(b12), P(SCR), n, N, st)
    e1.record()
    torch.cuda.synchronize()
    print(f'debug {flags:2d} ({what}): {e0.elapsed_time(e1) / 20:.4f} ms', flush=True)
L.azb_set_debug(0)
# wait-time breakdown of cluster 0, cycles per slab (probe build: nvcc ... -DAZB_PROF=1 -o tools/probe/libprof.so)
prof_lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libprof.so')
if not os.path.exists(prof_lib):
    sys.exit(0)
L = ctypes.CDLL(prof_lib)
L.az_nn_resblock.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p]
L.azb_set_prof.argtypes = [ctypes.c_void_p]
L.azb_set_prof.restype = None
L.azb_set_debug.argtypes = [ctypes.c_int]
prof = torch.zeros(2 * 8 * 8, dtype=torch.int64, device='cuda')
names = {'epi': ['wait staging tile', 'wait mma_done', 'wait residual', 'tmem + math + stores', '', ''],
         'mma': ['wait in_full', '', 'wait blk_free', 'wait turn', 'issue + commit + pass', ''],
         'ldr': ['wait stage free', '', '', '', '', ''],
         'rly': ['wait in_full', 'wait mma_done', '', '', '', ''], 'sto2': ['wait out_done', 'wait y_free (P)', '', '', '', '']}
for flags in (0, 6, 14):
    L.azb_set_debug(flags)
    L.azb_set_prof(P(prof))
    L.az_nn_resblock(P(x), P(w12), P(b12), P(SCR), n, N, st)
    torch.cuda.synchronize()
    L.azb_set_prof(None)
    t = prof.cpu().view(2, 8, 8)
    print(f'--- debug {flags}: cycles per slab, cluster 0')
    for rank, rn in ((0, 'P'), (1, 'C')):
        for role, (rname, kind) in enumerate((('epilogue g0', 'epi'), ('epilogue g1', 'epi'), ('mma even', 'mma'), ('mma odd', 'mma'), ('loader', 'ldr'), ('relay', 'rly'), ('storer even', 'sto2'), ('storer odd', 'sto2'))):
            row = t[rank, role]
            ns = max(int(row[7]), 1)
            per = ns if kind in ('ldr', 'rly', 'sto') else ns / 2      # these roles take every other slab
            items = ', '.join(f'{nm} {int(row[k]) / per:.0f}' for k, nm in enumerate(names[kind]) if nm)
            print(f'  {rn} {rname:12s} total {int(row[6]) / ns:.0f}/slab | per own slab: {items}')
L.azb_set_debug(0)
