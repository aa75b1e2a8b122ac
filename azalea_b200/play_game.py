"""One game between agents (azalea/play_game.py:18-139), unchanged API.

This is the single-game driver the reference's CLIs and tournament use; the
many-games self-play path is ``azalea_b200.selfplay``.
"""
import logging
import time
from collections import defaultdict
from typing import Dict, Sequence, Tuple

import numpy as np

from .replay_buffer import ReplayDataFrame


class _Wrapper:
    def __init__(self, agent):
        self._agent = agent

    def __getattr__(self, name):
        return getattr(self._agent, name)


class _Collect(_Wrapper):
    """play_game.py:81-98: record (state before move, pi) per ply."""

    def __init__(self, agent, game_data):
        super().__init__(agent)
        self._data = game_data

    def choose_action(self) -> int:
        move = self._agent.choose_action()
        self._data.state.append(self._agent.game.state)
        self._data.moves_prob.append(
            self._agent.info['moves_prob'].astype(np.float32))
        return move


class _Metrics(_Wrapper):
    """play_game.py:101-117."""

    def __init__(self, agent, metrics):
        super().__init__(agent)
        self._metrics = metrics

    def choose_action(self) -> int:
        move = self._agent.choose_action()
        for name, v in self._agent.info['metrics'].items():
            self._metrics[name] += v
        self._metrics['action_logprob'] += np.log(self._agent.info['prob'])
        return move


class _Print(_Wrapper):
    """play_game.py:120-139."""

    def choose_action(self) -> int:
        move = self._agent.choose_action()
        st = self._agent.game.state
        n = st.board.shape[0]
        color = ['white', 'black'][st.color]
        col, row = (move - 1) % n, (move - 1) // n
        print(f'Move {self._agent.ply + 1} ({color}): '
              f'{chr(ord("a") + col)}{row + 1} {self._agent.info["prob"]:.2f}')
        return move


def play_game(agents: Sequence, *, game_max_length: int = 300,
              print_moves: bool = False, collect_data: bool = False) \
        -> Tuple[int, ReplayDataFrame, Dict]:
    """Play one game; may raise SearchTreeFull.
    :returns: result, replay data, search metrics"""
    for a in agents:
        a.reset()
    game_data = ReplayDataFrame()
    if collect_data:
        agents = [_Collect(a, game_data) for a in agents]
    metrics: Dict[str, float] = defaultdict(int)
    agents = [_Metrics(a, metrics) for a in agents]
    if print_moves:
        agents = [_Print(a) for a in agents]

    start_time = time.time()
    result = 0
    for ply in range(game_max_length):
        move = agents[0].choose_action()
        res = [ag.execute_action(move) for ag in agents]
        result = res[0]
        assert all(r == result for r in res), 'conflicting game states'
        if result:
            break
        agents = agents[::-1]       # switch turns
    game_length = ply + 1
    assert result in (0, 1, 2, 3)
    if not result:
        logging.warning("game didn't terminate in %d moves", game_max_length)
        result = 2
    # play_game.py:63-67
    reward = np.full(game_length, result - 2., dtype=np.float32)
    reward[1::2] *= -1
    if collect_data:
        game_data.reward = list(reward)
    for name in metrics:
        metrics[name] /= max(1, game_length)
    metrics['games'] = 1
    metrics['reward'] = float(reward[-1])
    metrics['moves_per_game'] = game_length
    metrics['seconds_per_game'] = time.time() - start_time
    metrics['game_error'] = 0
    return result, game_data, metrics
