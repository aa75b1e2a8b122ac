"""GPU tests of the tournament driver (azalea/evaluation.py:17-80 semantics):
two trees per game, per-policy evaluator routing, outcome bookkeeping."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def stub_agent(n, sims, mode=2, seed=0):
    import azalea_b200 as az
    p = az.Policy()
    p.net = az.StubEvaluator(mode)
    p.simulations, p.search_batch_size, p.exploration_coef = sims, 6, 0.5
    p.exploration_depth, p.exploration_temperature = 4, 1.0
    p.exploration_noise_alpha, p.exploration_noise_scale = 0.03, 0.25
    a = az.AzaleaAgent(lambda: az.HexGame(n), policy=p)
    a.settings['move_sampling'] = True
    return a


def net_agent(n, seed, sims=30):
    import azalea_b200 as az
    torch.manual_seed(seed)
    p = az.Policy()
    p.initialize(dict(device='cuda', network='HexNetwork', board_size=n,
                      num_blocks=1, base_chans=32, simulations=sims,
                      search_batch_size=5, exploration_coef=0.5,
                      exploration_depth=4, exploration_noise_alpha=0.03,
                      exploration_noise_scale=0.25, exploration_temperature=1.0))
    a = az.AzaleaAgent(lambda: az.HexGame(n), policy=p)
    a.settings['move_sampling'] = True
    return a


def test_matches_are_legal_games_with_correct_results():
    """Every game of play_matches replays through the oracle: all moves
    legal, alternating, and the reported winner is the oracle's."""
    import azalea_b200 as az
    n, G = 5, 16 + 64 + 64
    a, b = stub_agent(n, 120).policy, stub_agent(n, 100, mode=1).policy
    rnd = az.RandomPolicy()
    first = [a] * 16 + [b] * 64 + [rnd] * 64
    second = [b] * 16 + [rnd] * 64 + [a] * 64
    result, history = az.play_matches(first, second, n, seed=3)
    assert set(result.tolist()) <= {1, 3}
    for g in range(G):
        game = oracle.Hex(n)
        for ply in range(len(history)):
            mv = int(history[ply, g])
            if mv == 0:
                break
            assert mv in game.legal_moves()
            game.step(mv)
        assert game.result() == result[g]
    # even with a noise evaluator the search sees terminal positions, so the
    # searching side beats the random mover on balance (first = b vs random:
    # b moves first; first = random vs a: a moves second)
    wins = (result[16:80] == 3).sum() + (result[80:144] == 1).sum()
    assert wins > 0.55 * 128, wins


def test_evaluate_round_robin_bookkeeping():
    """evaluate(agents, rounds) -> {pair: [wins_i, draws, wins_j]} over all
    pairs i < j (evaluation.py:17-46), random anchor first as compare() does
    (compare_cli.py:57-82); two different networks are routed to their own
    trees."""
    import azalea_b200 as az
    n, rounds = 5, 12
    agents = [az.AzaleaAgent(lambda: az.HexGame(n)),      # RandomPolicy
              net_agent(n, 1), net_agent(n, 2), stub_agent(n, 50)]
    out = az.evaluate(agents, rounds)
    assert sorted(out) == [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    for pair, (w, d, l) in out.items():
        assert w + d + l == rounds and d == 0      # Hex has no draws
    # the anchor loses to every searching policy on balance
    lost = sum(out[(0, j)][2] for j in (1, 2, 3))
    assert lost > 0.5 * 3 * rounds, out
    # first-move coin flips are the reference's: RandomState(10000*r + s)
    orders = [np.random.RandomState(10000 * r + s).choice([-1, 1])
              for r in range(rounds) for s in range(6)]
    assert 0.25 < np.mean(np.array(orders) == 1) < 0.75
