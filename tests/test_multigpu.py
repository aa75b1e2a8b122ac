"""Two-GPU test of the path's only exchange over NCCL: the replay gather to
rank 0 (selfplay.gather_replay_rows / LockstepSelfPlay.harvest(gather=True),
SURVEY 8e).  Needs two visible GPUs (`gpurun --gpus 2`); skipped otherwise.
The content of the gathered rows is checked against the same games played by
one process (world-size invariance), so a wrong count, a wrong offset or a
lost block shows up as a byte difference."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _selfplay_rows(rank, world, G, moves, gather):
    from azalea_b200 import LockstepSelfPlay, StubEvaluator
    sp = LockstepSelfPlay(StubEvaluator(2), num_games=G, board_size=5, simulations=40,
                          search_batch_size=4, exploration_depth=6, seed=17, rank=rank,
                          world_size=world, cuda_graph=False, move_exploration=False)
    chunks = []
    for _ in range(moves):
        sp.step_move()
        chunks.append(sp.harvest(gather=gather))     # every rank calls it every move
    return np.concatenate(chunks)


def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method=f'tcp://127.0.0.1:{port}', rank=rank,
                            world_size=world, device_id=torch.device('cuda', rank))
    rows = _selfplay_rows(rank, world, 48, 40, True)
    q.put((rank, rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_replay_gather_over_nccl_has_the_right_rows():
    import torch.multiprocessing as mp
    from azalea_b200.engine import decode_replay_rows
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(got[1]) == 0 and len(got[0]) > 0          # everything lands on rank 0
    # the same 96 slots played by one process
    whole = _selfplay_rows(0, 1, 96, 40, False)

    def by_game(rows):
        h, _, _ = decode_replay_rows(rows, 5)
        out = {}
        for i in range(len(h)):
            out.setdefault(int(h['game_id'][i]), []).append(rows[i].tobytes())
        return out
    a, b = by_game(got[0]), by_game(whole)
    first = [gid for gid in b if gid < 96]               # first game of every slot
    assert len(first) == 96
    ranks_seen = set()
    for gid in first:
        assert a[gid] == b[gid], gid
        ranks_seen.add(gid // 48)
    assert ranks_seen == {0, 1}
