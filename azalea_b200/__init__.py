"""azalea_b200: B200-native self-play search for Azalea (Hex MCTS).

Drop-in for the reference's hot path: ``AzaleaAgent`` / ``Policy`` /
``SearchTree`` / ``HexGame`` keep the reference's Python API, the search
itself runs in hand-written sm_100a kernels behind a C ABI
(include/azalea_b200.h), and ``LockstepSelfPlay`` / ``Player`` run thousands
of games per GPU.  There is no CPU fallback.
"""
from . import typing, utils
from .azalea_agent import AzaleaAgent
from .engine import Engine
from .evaluation import evaluate, play_matches
from .game.hex import HexGame
from .parallel_player import Player
from .play_game import play_game
from .policy import Policy
from . import policy_trainer
from .random_policy import RandomPolicy
from .replay_buffer import ReplayDataFrame, ReplayRecord
from .replay_device import DeviceReplayBuffer
from .search_tree import SearchTree, SearchTreeFull, as_distribution
from .selfplay import LockstepSelfPlay, StubEvaluator

__version__ = '0.1.0'
__all__ = ['AzaleaAgent', 'Engine', 'HexGame', 'Player', 'play_game',
           'Policy', 'RandomPolicy', 'ReplayDataFrame', 'ReplayRecord',
           'SearchTree', 'SearchTreeFull', 'as_distribution',
           'LockstepSelfPlay', 'StubEvaluator', 'DeviceReplayBuffer', 'evaluate',
           'play_matches', 'policy_trainer', 'typing', 'utils']
