"""Two halves of the games on two streams (tree kernels of one half under the network of the
other) against one lockstep instance: ms per move of all games."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from azalea_b200 import LockstepSelfPlay
from azalea_b200.network import HexNetwork
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
net = HexNetwork(11, 6, 64).eval().cuda()
net.prepare_inference(torch.bfloat16)

def time_it(step, n=4):
    for _ in range(3): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

one = LockstepSelfPlay(net, num_games=G, board_size=11, seed=1, collect_replay=False)
for _ in range(3): one.step_move()
print(f'one instance, {G} games: {time_it(one.step_move):.1f} ms per move', flush=True)
del one
torch.cuda.empty_cache()

for parts in (2, 3):
    halves = [LockstepSelfPlay(net, num_games=G // parts, board_size=11, seed=1 + i, collect_replay=False,
                               cuda_graph=False, rank=i, world_size=parts) for i in range(parts)]
    side = [torch.cuda.Stream() for _ in range(parts - 1)]
    def body():
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event(); fork.record(main)
        joins = []
        for s, h in zip(side, halves[1:]):
            s.wait_event(fork)
            with torch.cuda.stream(s):
                h._move_body()
                j = torch.cuda.Event(); j.record(s); joins.append(j)
        halves[0]._move_body()
        for j in joins: main.wait_event(j)
    for _ in range(2): body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    print(f'{parts} parts of {G // parts} games on {parts} streams: {time_it(g.replay):.1f} ms per move', flush=True)
    del halves, g
    torch.cuda.empty_cache()
