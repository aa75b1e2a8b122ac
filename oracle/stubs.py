"""Deterministic stub evaluators (numpy).  TEST INFRASTRUCTURE ONLY.

The same arithmetic exists in three places and must agree bit for bit:
``ostub_eval`` in azalea_oracle.c, ``az_stub_eval`` in the CUDA engine, and
this file.  This copy plugs into the *Python reference* in two ways:

* ``StubNet``: an object with the evaluator interface the reference search
  calls (``.eval()``, ``.device``, ``.run(batch)``; mcts.py:203-210,
  network.py:87-105).  Priors go through ``log`` here and ``np.exp`` on the
  reference's host side, exactly like a real network's would.
* ``make_evaluate_batch(mode)``: a drop-in for ``mcts.evaluate_batch``
  (mcts.py:155-217, looked up as a module global at mcts.py:25,282) that
  returns exactly rounded priors ``w/sum(w)`` so that visit counts can be
  compared bit for bit with implementations that never take a log.
"""
import numpy as np

UNIFORM, DYADIC, ROUGH = 0, 1, 2

_M32 = np.uint64(0xFFFFFFFF)


def fmix32(h):
    h = np.asarray(h, dtype=np.uint64) & _M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & _M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & _M32
    h ^= h >> np.uint64(16)
    return h


def board_hash(board_view):
    """Order-independent hash of one board in first-player view."""
    b = np.asarray(board_view).astype(np.int64).ravel()
    n = int(round(np.sqrt(b.size)))
    idx = np.flatnonzero(b)
    x = ((2 * idx + b[idx]).astype(np.uint64) * np.uint64(0x9E3779B1)) & _M32
    s = np.uint64(int(fmix32(x).sum()) & 0xFFFFFFFF)
    return int(fmix32(s ^ np.uint64(n)))


def stub_eval(mode, board_view, moves_view):
    """Value (float32) and exactly rounded priors (float32[k])."""
    moves_view = np.asarray(moves_view)
    moves_view = moves_view[moves_view > 0]
    h0 = board_hash(board_view)
    if mode == UNIFORM:
        value = np.float32(0.0)
    elif mode == DYADIC:
        value = np.float32((int((h0 >> 8) % 17) - 8)) / np.float32(8.0)
    else:
        value = np.float32(h0 >> 8) * np.float32(1.0 / 8388608.0) \
            - np.float32(1.0)
    t = (moves_view.astype(np.uint64) - np.uint64(1))
    hj = fmix32(np.uint64(h0) ^ ((t * np.uint64(0x9E3779B1)
                                  + np.uint64(0x7F4A7C15)) & _M32))
    if mode == UNIFORM:
        w = np.ones(len(moves_view), dtype=np.uint64)
    elif mode == DYADIC:
        w = np.uint64(1) + (hj >> np.uint64(28))
    else:
        w = np.uint64(1) + (hj >> np.uint64(24))
    prior = w.astype(np.float32) / np.float32(int(w.sum()))
    return np.float32(value), prior.astype(np.float32)


def make_evaluate_batch(mode):
    """Replacement for the reference's ``mcts.evaluate_batch``."""
    from azalea import prep  # the reference package (oracle use only)

    def evaluate_batch(game, net, states, rng):
        batch = prep.batch_games(states)
        batch['board'] = batch['board'].copy()
        batch['legal_moves'] = batch['legal_moves'].copy()
        flip = batch['color'] == 1
        batch['board'][flip], batch['legal_moves'][flip] = \
            game.flip_player_board_moves(batch['board'][flip],
                                         batch['legal_moves'][flip])
        value = np.zeros(len(states), dtype=np.float32)
        prior = np.zeros(batch['legal_moves'].shape, dtype=np.float32)
        num_children = (batch['legal_moves'] > 0).sum(1)
        for i in range(len(states)):
            if batch['result'][i] != 0:
                value[i] = -1.0
                continue
            v, p = stub_eval(mode, batch['board'][i], batch['legal_moves'][i])
            value[i] = v
            prior[i, :len(p)] = p
        return value, num_children, prior

    return evaluate_batch


class StubNet:
    """Evaluator object for the reference's (and the drop-in's) ``net`` slot.

    ``run`` returns ``moves_logprob = log(prior)`` as float32; the caller
    exponentiates on the host (mcts.py:210), so the priors that reach the
    tree are ``np.exp(np.log(w/sum w))`` on both sides of a comparison.
    """

    def __init__(self, mode=UNIFORM):
        import torch
        self.mode = mode
        self.device = torch.device('cpu')
        self.calls = 0
        self.rows = 0

    def eval(self):
        return self

    def run(self, batch):
        import torch
        board = batch['board'].cpu().numpy()
        moves = batch['legal_moves'].cpu().numpy()
        value = np.zeros(len(board), dtype=np.float32)
        logp = np.full(moves.shape, -99.0, dtype=np.float32)
        for i in range(len(board)):
            v, p = stub_eval(self.mode, board[i], moves[i])
            value[i] = v
            logp[i, :len(p)] = np.log(p)
        self.calls += 1
        self.rows += len(board)
        return dict(value=torch.from_numpy(value),
                    moves_logprob=torch.from_numpy(logp))
