"""Long self-play soak: every game slot plays several whole games; checks
status flags, counters and (for a sample) replays rows through the oracle."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from azalea_b200 import LockstepSelfPlay, StubEvaluator
from azalea_b200.engine import decode_replay_rows
from azalea_b200.network import HexNetwork

ap = argparse.ArgumentParser()
ap.add_argument('--evaluator', default='stub')
ap.add_argument('--games', type=int, default=4096)
ap.add_argument('--board', type=int, default=11)
ap.add_argument('--moves', type=int, default=260)
ap.add_argument('--streams', type=int, default=1)
ap.add_argument('--seconds', type=float, default=0, help='run for this long instead of --moves')
a = ap.parse_args()
if a.evaluator == 'net':
    torch.manual_seed(0); ev = HexNetwork(a.board, 6, 64).eval().cuda()
else:
    ev = StubEvaluator(2)
sp = LockstepSelfPlay(ev, num_games=a.games, board_size=a.board, simulations=800,
                      search_batch_size=10, exploration_coef=0.5, seed=3, streams=a.streams)
rows = []
t0 = time.time()
m = 0
while (time.time() - t0 < a.seconds) if a.seconds else (m < a.moves):
    sp.step_move()
    if m % 8 == 7 and sp.eng.replay_count():
        r = sp.harvest()
        if sum(len(x) for x in rows) < 1_500_000:   # keep a bounded sample for the oracle check
            rows.append(r)
    m += 1
a.moves = m
torch.cuda.synchronize()
dt = time.time() - t0
if sp.eng.replay_count():
    rows.append(sp.harvest())
rows = np.concatenate(rows)
cnt = sp.counters()
st = sp.eng.status().cpu().numpy()
print('seconds', dt, 'moves/s', a.games * a.moves / dt, 'sims/s', a.games * a.moves * 810 / dt)
print('counters', cnt)
print('status nonzero', int((st != 0).sum()), 'rows', len(rows))
h, board, visits = decode_replay_rows(rows, a.board)
games = {}
for i in range(len(h)):
    games.setdefault(int(h['game_id'][i]), []).append(i)
lens = np.array([len(v) for v in games.values()])
print('finished games', len(games), 'mean length', lens.mean(), 'min', lens.min(), 'max', lens.max())
bad = 0
for gid in list(games)[:300]:
    idx = games[gid]
    g = oracle.Hex(a.board)
    for r, i in enumerate(idx):
        assert h['ply'][i] == r and (board[i] == g.board).all()
        g.step(int(h['move'][i]))
    if g.result() != h['result'][idx[0]]:
        bad += 1
print('oracle-checked 300 games, mismatches', bad)
assert bad == 0 and (st == 0).all() and cnt['games_failed'] == 0 and cnt['replay_dropped'] == 0
print('useful NN rows fraction', cnt['nn_rows'] / cnt['simulations'])
