"""Round-robin tournament on the GPU (the compare / evaluation path).

Same call and result as the reference's ``evaluate(agents, num_rounds,
num_workers)`` (azalea/evaluation.py:17-80): every pair ``(i, j), i < j``
plays ``num_rounds`` games, the first mover of each game is a coin flip from
``RandomState(10000 * round + s)``, and the result is ``{pair: [wins of i,
draws, wins of j]}``.  The reference plays one game per worker process with
two agents, two trees and two networks, and pushes every move into both trees
(azalea/play_game.py:46-54).  Here all ``pairs x rounds`` games run in
lockstep: one engine holds the first movers' trees, another the second
movers'; at each ply the mover's engine searches, commits, and the move is
replayed into the other engine (``tree_move`` + ``hex_step``).  Leaves are
routed to the network of the policy that owns the tree.

Moves are drawn on the device (per-game Philox streams) instead of from each
agent's NumPy ``RandomState``: outcomes are statistically, not bitwise,
comparable with the reference's.  ``RandomPolicy`` agents move uniformly at
random without search (azalea/random_policy.py:25-41).
"""
from collections import defaultdict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _cabi
from .engine import Engine
from .random_policy import RandomPolicy
from .selfplay import StubEvaluator

Pair = Tuple[int, int]
M_STATUS = 5


def gen_pairs(num_players: int) -> List[Pair]:
    """Round robin pair ordering (evaluation.py:41-46)."""
    return [(i, j) for j in range(num_players) for i in range(j)]


def _policy_key(policy):
    if isinstance(policy, RandomPolicy):
        return ('random',)
    net = policy.net
    kind = ('stub', net.mode) if isinstance(net, StubEvaluator) else ('net',)
    sampling = bool(policy.settings.get('move_sampling', False))
    noise = policy.exploration_noise_scale \
        if (sampling and policy.settings.get('move_exploration', False)) else 0.0
    return (policy.simulations, policy.search_batch_size,
            float(policy.exploration_coef), float(policy.exploration_temperature),
            int(policy.exploration_depth), sampling, float(noise),
            float(policy.exploration_noise_alpha)) + kind


class _Seat:
    """One engine = the trees of all games' first (or second) movers."""

    def __init__(self, policies, board_size, seed, device, nodes_per_game):
        self.policies = policies                # policy object per game
        G = len(policies)
        batch = max([p.search_batch_size for p in policies
                     if not isinstance(p, RandomPolicy)] + [1])
        sims = max([p.simulations for p in policies
                    if not isinstance(p, RandomPolicy)] + [1])
        if nodes_per_game is None:
            nodes_per_game = 2 * (sims + batch + 1) * board_size ** 2
        self.eng = Engine(G, board_size, max_batch=batch,
                          nodes_per_game=nodes_per_game, seed=seed,
                          device=device)
        self.device = self.eng.device
        self.chosen = torch.zeros(G, 4, dtype=torch.int32, device=self.device)
        # games grouped by search configuration, then by evaluator object
        self.groups = defaultdict(lambda: defaultdict(list))
        for g, p in enumerate(policies):
            key = _policy_key(p)
            self.groups[key][id(None if key[0] == 'random' else p.net)].append(g)
        self.nets = {id(p.net): p.net for p in policies
                     if not isinstance(p, RandomPolicy)}
        for net in self.nets.values():
            if not isinstance(net, StubEvaluator):
                net.eval()
                net.to(self.device)
                if getattr(net, '_fast', None) is None:
                    net.prepare_inference()
        self.index = {k: {nid: torch.tensor(gs, device=self.device)
                          for nid, gs in by_net.items()}
                      for k, by_net in self.groups.items()}
        self.rng = torch.Generator(device=self.device)
        self.rng.manual_seed(seed)

    def _pause_all_but(self, games):
        pause = torch.full((self.eng.num_games,), _cabi.AZ_ST_DISABLED,
                           dtype=torch.int32, device=self.device)
        pause[games] = 0
        status = self.eng.meta[:, M_STATUS]
        self.eng.meta[:, M_STATUS] = (status & ~_cabi.AZ_ST_DISABLED) | pause

    def _resume_all(self):
        status = self.eng.meta[:, M_STATUS]
        self.eng.meta[:, M_STATUS] = status & ~_cabi.AZ_ST_DISABLED

    def _evaluate(self, key, by_net, root):
        eng = self.eng
        if key[-2] == 'stub':
            eng.stub_eval(key[-1])
            return _cabi.AZ_PRIOR_PROBS
        B = 1 if root else eng.max_batch
        for nid, idx in by_net.items():
            cells = eng.leaf_board[idx, :B].reshape(-1, eng.cell_stride)
            value, logits = self.nets[nid].evaluate_cells(cells)
            eng.value[idx, :B] = value.view(len(idx), B)
            eng.prior[idx, :B] = logits.view(len(idx), B, eng.nn)
        return _cabi.AZ_PRIOR_LOGITS

    def play_ply(self):
        """Every unfinished game's mover (this seat) searches and moves.
        Returns (move, move_id) int32[G] tensors (0 / -1 for finished games)."""
        eng = self.eng
        self.chosen.zero_()
        self.chosen[:, 1] = -1
        for key, by_net in self.index.items():
            games = torch.cat(list(by_net.values()))
            self._pause_all_but(games)
            if key[0] == 'random':
                self._random_moves(games)
                continue
            sims, batch, coef, temp, depth, sampling, noise, alpha = key[:8]
            eng.select_root()
            kind = self._evaluate(key, by_net, True)
            eng.expand_root(None, kind)
            for _ in range(sims // batch + 1):
                eng.select(batch, coef, noise, alpha)
                kind = self._evaluate(key, by_net, False)
                eng.expand_backup(None, None, kind)
            eng.play_commit(temp, depth, sampling, False, False, self.chosen)
        self._resume_all()
        return self.chosen[:, 0].clone(), self.chosen[:, 1].clone()

    def _random_moves(self, games):
        """RandomPolicy.choose_action (random_policy.py:25-41) for `games`."""
        eng = self.eng
        moves, count = eng.hex_legal_moves()
        count = count[games]
        live = count > 0
        u = torch.rand(len(games), device=self.device, generator=self.rng)
        ordinal = torch.minimum((u * count).long(), (count - 1).clamp(min=0).long())
        picked = moves[games].gather(1, ordinal[:, None]).squeeze(1)
        step = torch.zeros(eng.num_games, dtype=torch.int32, device=self.device)
        step[games] = torch.where(live, picked, torch.zeros_like(picked))
        ids = -torch.ones(eng.num_games, dtype=torch.int32, device=self.device)
        ids[games] = torch.where(live, ordinal.int(), -torch.ones_like(ordinal).int())
        self._resume_all()
        eng.tree_move(ids)
        eng.hex_step(step)
        self.chosen[games, 0] = step[games]
        self.chosen[games, 1] = ids[games]

    def apply_opponent(self, moves, move_ids):
        """AzaleaAgent.execute_action for the opponent's move
        (azalea_agent.py:60-64): tree first, then game."""
        self.eng.tree_move(move_ids)
        self.eng.hex_step(moves)


def play_matches(first: List, second: List, board_size: int, seed: int = 0,
                 device=None, nodes_per_game: Optional[int] = None,
                 game_max_length: int = 300):
    """Play len(first) games in lockstep; game g is first[g] (moves first)
    against second[g].  Returns (results int array: 3 first mover won, 1
    second mover won, 2 draw; move history int32 [plies, G])."""
    assert len(first) == len(second)
    seats = [_Seat(first, board_size, 2 * seed + 1, device, nodes_per_game),
             _Seat(second, board_size, 2 * seed + 2, device, nodes_per_game)]
    history = []
    G = len(first)
    result = np.zeros(G, dtype=np.int64)
    for ply in range(min(game_max_length, board_size ** 2)):
        mover, other = seats[ply % 2], seats[1 - ply % 2]
        moves, move_ids = mover.play_ply()
        other.apply_opponent(moves, move_ids)
        history.append(moves.cpu().numpy())
        res = mover.eng.hex_state()[2].cpu().numpy()
        result = res
        bad = (mover.eng.status() | other.eng.status()).cpu().numpy()
        assert not (bad & _cabi.AZ_ST_ILLEGAL).any(), 'inconsistent game state'
        if (res != 0).all():
            break
    result = np.where(result == 0, 2, result)       # play_game.py:57-61
    return result, np.stack(history)


def evaluate(agents: List, num_rounds: int, num_workers: Optional[int] = None,
             device=None) -> Dict[Pair, List[int]]:
    """Round robin tournament between agents (evaluation.py:17-38).
    ``num_workers`` is accepted for compatibility and ignored."""
    pairs = gen_pairs(len(agents))
    board_size = agents[0].game.board_size
    first, second, meta = [], [], []
    for r in range(num_rounds):
        for s, pair in enumerate(pairs):
            # evaluation.py:67-76: coin flip for the first move
            rng = np.random.RandomState(10000 * r + s)
            order = rng.choice([-1, 1])
            pa, pb = agents[pair[0]].policy, agents[pair[1]].policy
            if order == 1:
                first.append(pa)
                second.append(pb)
            else:
                first.append(pb)
                second.append(pa)
            meta.append((pair, order))
    result, _ = play_matches(first, second, board_size,
                             seed=int(np.random.RandomState(num_rounds).randint(1 << 30)),
                             device=device)
    outcomes: Dict[Pair, List[int]] = defaultdict(lambda: [0, 0, 0])
    for (pair, order), res in zip(meta, result):
        outcome = order * (int(res) - 2)
        outcomes[pair][0] += outcome > 0
        outcomes[pair][1] += outcome == 0
        outcomes[pair][2] += outcome < 0
    return dict(outcomes)
