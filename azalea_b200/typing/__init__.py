"""Boundary types of the hot path (reference: azalea/typing/agent.py:7-41,
azalea/typing/searchable_env.py:8-45)."""
from dataclasses import dataclass
from enum import IntEnum
from typing import Any, Mapping, Optional, Protocol

import numpy as np


class GameResult(IntEnum):
    """Wins/losses from the first player's perspective (agent.py:34-41)."""
    ONGOING = 0
    LOSS = 1
    DRAW = 2
    WIN = 3


@dataclass
class GameState:
    """searchable_env.py:41-45 / hex.py:11-16."""
    color: int              # 0 = first player to move, 1 = second
    legal_moves: np.ndarray  # int32, 1-based tile ids, ascending
    result: int
    board: np.ndarray       # int32 [n, n]: 0 empty, 1 X, 2 O


class Agent(Protocol):
    def reset(self, *args, **kwargs) -> None: ...
    def seed(self, seed: Optional[int]) -> None: ...
    @property
    def settings(self) -> Mapping[str, Any]: ...
    def choose_action(self) -> int: ...
    def execute_action(self, action: int) -> 'GameResult': ...


class SearchableEnv(Protocol):
    def reset(self, *args, **kwargs) -> None: ...
    def seed(self, seed: Optional[int]) -> None: ...
    def step(self, action: int) -> None: ...
    @property
    def state(self) -> GameState: ...
    def snapshot(self) -> None: ...
    def restore(self) -> None: ...


__all__ = ['Agent', 'SearchableEnv', 'GameState', 'GameResult']
