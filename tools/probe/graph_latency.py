"""Probe: what a host sync per move costs the lockstep step.  K moves replayed back to back
(the launches hide behind the previous move) against K moves with a synchronize after each
one, and the CPU time of one graph launch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from azalea_b200 import LockstepSelfPlay
from azalea_b200.network import HexNetwork
torch.manual_seed(0)
net = HexNetwork(11, 6, 64).eval().cuda()
net.prepare_inference(torch.bfloat16)
sp = LockstepSelfPlay(net, num_games=4096, board_size=11, simulations=800, search_batch_size=10, seed=1)
for _ in range(4):
    sp.step_move()
torch.cuda.synchronize()
K = 5
t0 = time.perf_counter()
for _ in range(K):
    sp.step_move()
torch.cuda.synchronize()
a = (time.perf_counter() - t0) / K
t0 = time.perf_counter()
cpu = 0.0
for _ in range(K):
    c0 = time.perf_counter()
    sp.step_move()
    cpu += time.perf_counter() - c0
    torch.cuda.synchronize()
b = (time.perf_counter() - t0) / K
print(f'back to back {a * 1e3:.2f} ms/move; with a sync after every move {b * 1e3:.2f} ms/move; '
      f'graph launch call {cpu / K * 1e3:.2f} ms CPU')
t0 = time.perf_counter()
for _ in range(K):
    net.prepare_inference(torch.bfloat16)
c = (time.perf_counter() - t0) / K
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(K):
    net.prepare_inference(torch.bfloat16)
torch.cuda.synchronize()
d = (time.perf_counter() - t0) / K
print(f'prepare_inference: {c * 1e3:.2f} ms CPU enqueue, {d * 1e3:.2f} ms with the GPU idle otherwise')
