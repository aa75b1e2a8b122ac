"""Uniform random mover, no search (azalea/random_policy.py:6-50)."""
from typing import Dict, Optional

import numpy as np


class RandomPolicy:
    def __init__(self):
        self.rng = np.random.RandomState()
        self.settings = {}
        self.ply = 0
        self.seed()

    def reset(self):
        self.ply = 0

    def seed(self, seed: Optional[int] = None) -> None:
        self.rng.seed(seed)

    def load_state_dict(self, state: Dict) -> None:
        pass

    def state_dict(self) -> Dict:
        return {}

    def choose_action(self, game):
        state = game.state
        assert not state.result
        moves = state.legal_moves
        action = self.rng.randint(len(moves))
        probs = np.ones(len(moves), dtype=np.float32) / len(moves)
        info = dict(move_id=action, moves=moves, moves_prob=probs,
                    prob=probs[action], metrics={})
        return moves[action], info

    def execute_action(self, move, moves):
        self.ply += 1

    def tree_metrics(self):
        return {}
