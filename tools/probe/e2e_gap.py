"""Probe: what sits between two moves of the end-to-end loop (bench.py's e2e leg): stream time of the
weight upload (82 H2D copies), of prepare_inference's kernels, and of the harvest, each alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from azalea_b200 import LockstepSelfPlay
from azalea_b200.network import HexNetwork
torch.manual_seed(0)
net = HexNetwork(11, 6, 64).eval().cuda()
net.prepare_inference(torch.bfloat16)
host_w = [p.detach().float().cpu().pin_memory() for p in net.parameters()]
sp = LockstepSelfPlay(net, num_games=4096, board_size=11, simulations=800, search_batch_size=10, seed=1)
for _ in range(3):
    sp.step_move()
sp.preroll(110)
for _ in range(2):
    sp.step_move()
sp.harvest()
torch.cuda.synchronize()

def timed(fn, reps=5):
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sp.step_move()              # the GPU is busy while the host enqueues fn's work behind it
        a.record(); fn(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / reps

def copies():
    with torch.no_grad():
        for p, w in zip(net.parameters(), host_w):
            p.copy_(w, non_blocking=True)

print('82 H2D parameter copies: %.3f ms of stream time' % timed(copies))
print('prepare_inference: %.3f ms of stream time' % timed(lambda: net.prepare_inference(torch.bfloat16)))
t0 = time.perf_counter(); h = sp.harvest_begin(); t1 = time.perf_counter(); rows = sp.harvest_end(h); t2 = time.perf_counter()
print('harvest_begin %.3f ms host (%d rows), harvest_end %.3f ms host' % ((t1 - t0) * 1e3, len(rows), (t2 - t1) * 1e3))
