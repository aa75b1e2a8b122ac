"""Time the UNMODIFIED Python/Numba reference on this (build) box -- the figure
bench.py quotes beside its CPU-port baseline (the reference cannot travel to
the GPU box: /root/reference does not exist there).

Workload = BASELINE.json config 1 / BASELINE.md section 3: Hex 11x11,
AzaleaAgent(HexGame, policy=Policy) with a picklable uniform stub evaluator,
search parameters of config/hex11_train_config.yml (800 simulations, batch 10,
c_puct 0.5), (a) one process, steady per-move time, and (b) the reference's
own ProcessPool + Player.read on all cores with OMP/MKL threads = 1.

    python tools/measure_reference.py [/root/reference] -> profiles/r02_reference_python.json
"""
import json
import os
import sys
import time

os.environ.setdefault('OMP_NUM_THREADS', '1')
os.environ.setdefault('MKL_NUM_THREADS', '1')
REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

import numba                            # noqa: E402
import numba.experimental              # noqa: E402
numba.jitclass = numba.experimental.jitclass    # the reference imports the old name (hex.py:6)

import numpy as np                      # noqa: E402
import torch                            # noqa: E402
import azalea as az                     # noqa: E402
from azalea.game.hex import HexGame     # noqa: E402
from ref_stub import UniformStub        # noqa: E402  (module-level: picklable for spawn workers)


def game_factory():
    return HexGame(11)


def make_agent():
    policy = az.Policy()
    policy.net = UniformStub()
    policy.simulations, policy.search_batch_size = 800, 10
    policy.exploration_coef, policy.exploration_depth = 0.5, 15
    policy.exploration_noise_alpha, policy.exploration_noise_scale = 0.03, 0.25
    policy.exploration_temperature = 1.0
    agent = az.AzaleaAgent(game_factory, policy=policy)
    agent.settings['move_sampling'] = True
    agent.settings['move_exploration'] = True
    return agent


def main():
    torch.set_num_threads(1)
    out = {'box': 'build container', 'cores': os.cpu_count(), 'workload':
           'Hex 11x11, Policy + uniform stub evaluator, 800 sims batch 10 (810 descents/move), '
           'self-play settings (temperature sampling + Dirichlet noise)'}
    agent = make_agent()
    agent.reset()
    agent.seed(1)
    agent.execute_action(agent.choose_action())        # JIT warm-up
    t0 = time.time()
    moves = 0
    while time.time() - t0 < 20 and not agent.game.state.result:
        agent.execute_action(agent.choose_action())
        moves += 1
    dt = time.time() - t0
    out['one_process'] = {'moves': moves, 'seconds': dt, 'moves_per_s': moves / dt,
                          'sims_per_s': 810 * moves / dt, 's_per_move': dt / moves}
    print(out['one_process'], flush=True)
    # the reference's own data-parallel path: ProcessPool(spawn) + Player.read
    from azalea.process_pool import ProcessPool
    from azalea.parallel_player import Player
    workers = os.cpu_count() or 1
    os.environ['PYTHONPATH'] = os.pathsep.join(
        [os.path.join(ROOT, 'tools', 'refshim'), os.path.join(ROOT, 'tools'), REF,
         os.environ.get('PYTHONPATH', '')])
    pool = ProcessPool(num_workers=workers)
    player = Player(pool, [make_agent()])
    t0 = time.time()
    player.read(100 * workers // 2)                     # import + JIT + first games in every worker
    warm = time.time() - t0
    t0 = time.time()
    data, metrics = player.read(1200)
    dt = time.time() - t0
    n = len(data)
    out['process_pool'] = {'workers': workers, 'positions': n, 'seconds': dt, 'warmup_seconds': warm,
                           'moves_per_s': n / dt, 'sims_per_s': 810 * n / dt,
                           'note': 'Player.read returns whole games as they finish; positions of games '
                                   'already in flight at the start are counted, as the reference does'}
    print(out['process_pool'], flush=True)
    player.stop()
    try:
        pool.close()
    except Exception:
        pass
    with open(os.path.join(ROOT, 'profiles', 'r02_reference_python.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
