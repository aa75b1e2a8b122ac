"""Run a few eager (un-graphed) lockstep moves for ncu.

    ncu ... python tools/profile_step.py --evaluator stub --warm 4 --moves 1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from azalea_b200 import LockstepSelfPlay, StubEvaluator  # noqa: E402
from azalea_b200.network import HexNetwork  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--evaluator', default='stub')
ap.add_argument('--games', type=int, default=4096)
ap.add_argument('--board', type=int, default=11)
ap.add_argument('--warm', type=int, default=4)
ap.add_argument('--moves', type=int, default=1)
ap.add_argument('--noise', type=int, default=1)
a = ap.parse_args()
if a.evaluator == 'net':
    torch.manual_seed(0)
    ev = HexNetwork(a.board, 6, 64).eval().cuda()
else:
    ev = StubEvaluator(2)
sp = LockstepSelfPlay(ev, num_games=a.games, board_size=a.board, simulations=800,
                      search_batch_size=10, exploration_coef=0.5, seed=1,
                      cuda_graph=False, move_exploration=bool(a.noise))
for _ in range(a.warm + a.moves):
    sp.step_move()
torch.cuda.synchronize()
print(sp.counters())
