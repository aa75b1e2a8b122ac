"""Replay data frame: the wire format self-play hands to the trainer
(azalea/replay_buffer.py:11-104).  Lists of (GameState, moves_prob, reward)."""
from dataclasses import dataclass, field, fields
from typing import List

import numpy as np

from .typing import GameState


@dataclass
class ReplayRecord:
    state: GameState
    moves_prob: np.ndarray
    reward: np.float32


@dataclass
class ReplayDataFrame:
    state: List[GameState] = field(default_factory=list)
    moves_prob: List[np.ndarray] = field(default_factory=list)
    reward: List[np.float32] = field(default_factory=list)

    def __len__(self) -> int:
        return len(self.state)

    def __getitem__(self, idx):
        rows = [getattr(self, f.name)[idx] for f in fields(self)]
        if isinstance(idx, int):
            return ReplayRecord(*rows)
        if isinstance(idx, slice):
            return ReplayDataFrame(*rows)
        raise TypeError(idx)

    def __setitem__(self, idx, rows):
        before = len(self)
        for f in fields(self):
            getattr(self, f.name)[idx] = getattr(rows, f.name)
        assert len(self) == before

    def append(self, rows: 'ReplayDataFrame') -> None:
        for f in fields(self):
            getattr(self, f.name).extend(getattr(rows, f.name))
