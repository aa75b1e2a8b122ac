"""Micro-benchmark of evaluator formulations (library calls only)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from azalea_b200.network import HexNetwork

torch.backends.cudnn.benchmark = True
dev = torch.device('cuda')
torch.manual_seed(0)
net = HexNetwork(11, 6, 64).eval().to(dev)
net.prepare_inference(torch.bfloat16)
f = net._fast
N = 40960
cells = torch.randint(0, 3, (N, 128), dtype=torch.int8, device=dev)

def timeit(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def base(c=cells):
    return net.evaluate_cells(c)

def unfused(c=cells):
    """Separate conv / bias / add / relu kernels (what plain F.conv2d gives)."""
    n = 11; M = c.shape[0]
    idx = c[:, :121].to(torch.int32)
    x = F.embedding(idx, f['emb']).view(M, n, n, 4).permute(0, 3, 1, 2)
    x = F.relu_(F.conv2d(x, *f['stem'], padding=1))
    for (w1, b1), (w2, b2) in f['blocks']:
        y = F.relu_(F.conv2d(x, w1, b1, padding=1))
        y = F.conv2d(y, w2, b2, padding=1)
        x = F.relu_(y.add_(x))
    h = F.relu_(F.conv2d(x, *f['heads']))
    flat = h.permute(0, 2, 3, 1).reshape(M, -1)
    y = F.linear(flat, *f['fc'])
    k2 = f['nfc2']
    value = torch.tanh(F.linear(F.relu(y[:, :k2]), *f['value_fc3'])).squeeze(1)
    return value.float(), y[:, k2:].float()


def chunked(fn, chunk):
    def run():
        outs = [fn(cells[i:i + chunk]) for i in range(0, N, chunk)]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
    return run

flop = 107.852e6 * N
res = {}
v0, l0 = base()
for name, fn in [('evaluate_cells', base), ('unfused', unfused)]:
    try:
        v, l = fn()
        err = (v - v0).abs().max().item(), (l - l0).abs().max().item()
        ms = timeit(fn)
        print(f'{name:24s} {ms:8.3f} ms  {flop/ms/1e9:8.1f} TFLOP/s  err {err}', flush=True)
    except Exception as e:
        print(name, 'FAILED', repr(e)[:300], flush=True)
for chunk in (2048, 4096, 8192, 16384):
    for name, fn in [('evaluate_cells', base), ('unfused', unfused)]:
        try:
            run = chunked(fn, chunk)
            ms = timeit(run)
            print(f'{name}-chunk{chunk:<14d} {ms:8.3f} ms  {flop/ms/1e9:8.1f} TFLOP/s', flush=True)
        except Exception as e:
            print(name, chunk, 'FAILED', repr(e)[:300], flush=True)
# just the 12 tower convs without elementwise: upper bound of cuDNN conv itself
x = torch.randn(N, 64, 11, 11, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
w = f['blocks'][0][0][0]
ms = timeit(lambda: F.conv2d(x, w, None, padding=1))
print(f'single conv3x3 64->64 N={N}: {ms:.3f} ms {2*121*64*576*N/ms/1e9:.1f} TFLOP/s')
x4 = x[:4096].contiguous(memory_format=torch.channels_last)
ms = timeit(lambda: F.conv2d(x4, w, None, padding=1), reps=20)
print(f'single conv3x3 64->64 N=4096: {ms:.3f} ms {2*121*64*576*4096/ms/1e9:.1f} TFLOP/s')
# as explicit GEMM via unfold-free trick: 1x1 conv == matmul
a = torch.randn(N * 121, 576, device=dev, dtype=torch.bfloat16)
bmat = torch.randn(576, 64, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: a @ bmat)
print(f'GEMM [{N*121}x576]x[576x64]: {ms:.3f} ms {2*121*64*576*N/ms/1e9:.1f} TFLOP/s')
