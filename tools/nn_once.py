import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from azalea_b200.network import HexNetwork
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = HexNetwork(11, 6, 64).eval().cuda()
net.prepare_inference(torch.bfloat16)
cells = torch.randint(0, 3, (40960, 128), dtype=torch.int8, device='cuda')
for _ in range(4):
    net.evaluate_cells(cells)
torch.cuda.synchronize()
