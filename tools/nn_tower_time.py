import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from azalea_b200.network import HexNetwork
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = HexNetwork(11, 6, 64).eval().cuda()
net.prepare_inference(torch.bfloat16)
cells = torch.randint(0, 3, (40960, 128), dtype=torch.int8, device='cuda')
for tower in ('cudnn', 'tcgen05'):
    net.tower = tower
    for _ in range(3): net.evaluate_cells(cells)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): net.evaluate_cells(cells)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f'{tower:8s} evaluate_cells(40960 boards): {ms:.3f} ms  {107.852e6*40960/ms/1e9:.1f} TFLOP/s', flush=True)
