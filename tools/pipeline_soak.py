"""Two lockstep instances (separate engines) on two streams, graphed, for many moves:
does anything break when two evaluator launches are in flight?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from azalea_b200 import LockstepSelfPlay
from azalea_b200.network import HexNetwork
moves = int(sys.argv[1]) if len(sys.argv) > 1 else 40
torch.manual_seed(0)
net = HexNetwork(11, 6, 64).eval().cuda()
net.prepare_inference(torch.bfloat16)
print('tower', net.tower, flush=True)
halves = [LockstepSelfPlay(net, num_games=2048, board_size=11, seed=1 + i, collect_replay=False,
                           cuda_graph=False, rank=i, world_size=2) for i in range(2)]
side = torch.cuda.Stream()
def body():
    main = torch.cuda.current_stream()
    fork = torch.cuda.Event(); fork.record(main)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        halves[1]._move_body()
        j = torch.cuda.Event(); j.record(side)
    halves[0]._move_body()
    main.wait_event(j)
for _ in range(2): body()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    body()
t0 = time.time()
for m in range(moves):
    g.replay()
    if m % 10 == 9:
        torch.cuda.synchronize()
        print('move', m + 1, 'ok', f'{(time.time() - t0) / (m + 1) * 1e3:.1f} ms/move', flush=True)
torch.cuda.synchronize()
print('done', [h.counters()['games_failed'] for h in halves], flush=True)
