#!/usr/bin/env python
"""Benchmark of the self-play search hot path (BASELINE.json metric:
MCTS simulations/sec and self-play moves/sec, Hex 11x11).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One *step* = one lockstep move of every resident game: evaluate_root +
81 x (select, evaluate, expand/backup) + move commit = 810 root-to-leaf
descents per game (mcts.py:268).  Default workload = BASELINE.json
configs[1]: 4096 concurrent games per GPU, 6x64 resnet random-init in bf16,
self-play settings (temperature sampling + Dirichlet root noise), search
parameters of config/hex11_train_config.yml.  Games shard by index over
ranks (no collective on the path): scaling is weak.

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = 'mcts_simulations_per_sec'
UNIT = 'simulations/s'
SEARCH = dict(simulations=800, search_batch_size=10, exploration_coef=0.5,
              exploration_depth=15, exploration_noise_alpha=0.03,
              exploration_noise_scale=0.25, exploration_temperature=1.0)
NN_FLOP_PER_LEAF = {11: 107.852e6, 19: 322.466e6}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=('ours', 'reference'))
    ap.add_argument('--games', type=int, default=4096, help='games per GPU')
    ap.add_argument('--board', type=int, default=11)
    ap.add_argument('--evaluator', default='net', choices=('net', 'stub'))
    ap.add_argument('--stub-mode', type=int, default=2)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-games', type=int, default=0)
    ap.add_argument('--skip-tree-only', action='store_true')
    ap.add_argument('--sims', type=int, default=800, help='simulations per move')
    ap.add_argument('--nodes-per-game', type=int, default=0)
    ap.add_argument('--streams', type=int, default=1,
                    help='windows of the games driven on separate streams (opt-in, see DESIGN.md 5)')
    args = ap.parse_args()
    SEARCH['simulations'] = args.sims
    return args


def workload_name(args):
    ev = '6x64 resnet random-init bf16' if args.evaluator == 'net' \
        else f'stub evaluator mode {args.stub_mode}'
    return (f'Hex {args.board}x{args.board} lockstep self-play, {args.games} '
            f'concurrent games/GPU, {ev}, {args.sims} sims batch 10')


# ------------------------------------------------------------------ clocks --

class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(
                    ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                     '--format=csv,noheader,nounits'], capture_output=True,
                    text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap')
        reasons = [nm for i, nm in enumerate(names)
                   if any(r[3 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.rows[0][1]),
                'power_w_max': max(float(r[2]) for r in self.rows),
                'samples': len(self.rows), 'reasons': reasons}


# ---------------------------------------------------------- CPU baseline ----

def cpu_selfplay_sample(board, evaluator, stub_mode, games, moves, threads):
    """The oracle (C port of the reference's search + rules) timed on host
    cores on a bounded sample of the workload: `games` whole-tree searches
    advanced `moves` plies.  With the network evaluator the leaves of all
    sampled games are batched into one fp32 PyTorch-CPU forward per search
    batch (the reference evaluates <= 10 leaves per call; batching across
    games only helps the baseline)."""
    import oracle
    sims, batch, coef = SEARCH['simulations'], SEARCH['search_batch_size'], \
        SEARCH['exploration_coef']
    per_move = (sims // batch + 1) * batch
    if evaluator == 'stub':
        t0 = time.time()
        # whole games on C threads; bounded by the number of games
        r = oracle.bench_selfplay_stub(n=board, num_games=games, threads=threads,
                                       num_simulations=sims, batch_size=batch,
                                       coef=coef, stub_mode=stub_mode,
                                       exploration_depth=SEARCH['exploration_depth'],
                                       max_nodes=10_000_000, seed=1)
        secs = r['seconds']
        return dict(sims=r['simulations'], plies=r['plies'], seconds=secs,
                    sample=f'{games} whole games, {r["plies"]} plies, '
                           f'{threads} C threads, stub evaluator')
    import torch
    from azalea_b200.network import HexNetwork
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = HexNetwork(board, 6, 64).eval()
    nn = board * board
    hexes = [oracle.Hex(board) for _ in range(games)]
    trees = [oracle.Tree(max_nodes=2_000_000) for _ in range(games)]
    rng = np.random.RandomState(0)

    def evaluate(all_leaves):
        rows = [(g, i) for g, lv in enumerate(all_leaves)
                for i, leaf in enumerate(lv) if not leaf['result']]
        value = [np.zeros(len(lv), np.float32) for lv in all_leaves]
        prior = [np.zeros((len(lv), nn), np.float32) for lv in all_leaves]
        if rows:
            boards = np.stack([all_leaves[g][i]['board'] for g, i in rows])
            K = max(len(all_leaves[g][i]['legal_moves']) for g, i in rows)
            moves_ = np.zeros((len(rows), K), np.int32)
            for r, (g, i) in enumerate(rows):
                lm = all_leaves[g][i]['legal_moves']
                moves_[r, :len(lm)] = lm
            with torch.no_grad():
                out = net.run(dict(board=torch.from_numpy(boards),
                                   legal_moves=torch.from_numpy(moves_)))
            v = out['value'].numpy()
            p = np.exp(out['moves_logprob'].numpy())
            for r, (g, i) in enumerate(rows):
                value[g][i] = v[r]
                prior[g][i, :K] = p[r]
        return value, prior

    def one_move():
        need = [g for g in range(games) if not trees[g].root_evaluated()]
        if need:
            leaves = [trees[g].root_leaf(hexes[g]) for g in need]
            _, prior = evaluate(leaves)
            for j, g in enumerate(need):
                trees[g].expand_root(prior[j][0])
        for _ in range(sims // batch + 1):
            # select for all games, one network call, expand/backup for all
            leaves = [trees[g].select_batch(hexes[g], batch, coef)
                      for g in range(games)]
            value, prior = evaluate(leaves)
            for g in range(games):
                trees[g].expand_backup(value[g], prior[g])
        for g in range(games):
            v, _, _ = trees[g].root_stats()
            mid = int(rng.choice(len(v), p=v / v.sum()))
            mv = hexes[g].legal_moves()[mid]
            trees[g].move(mid)
            hexes[g].step(int(mv))
            if hexes[g].result():
                hexes[g] = oracle.Hex(board)
                trees[g].reset()

    one_move()      # warm-up (allocations, torch thread pool)
    t0 = time.time()
    for _ in range(moves):
        one_move()
    secs = time.time() - t0
    return dict(sims=games * moves * per_move, plies=games * moves, seconds=secs,
                sample=f'{games} games x {moves} plies, C oracle tree + fp32 '
                       f'PyTorch-CPU 6x64 net ({threads} threads), leaves of all '
                       f'sampled games batched per search batch')


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path.
    The Python/Numba reference does not travel to the GPU box, so this is the
    oracle port (kind = 'port'), on all host cores."""
    cores = os.cpu_count() or 1
    threads = max(1, cores)
    games = args.cpu_games or (4 * threads if args.evaluator == 'net'
                               else 16 * threads)
    tot_sims = tot_plies = 0
    tot_secs = 0.0
    sample = ''
    for step in range(args.warmup + args.steps):
        if args.evaluator == 'net':
            r = cpu_selfplay_sample(args.board, 'net', args.stub_mode, games, 1, threads)
        else:
            r = cpu_selfplay_sample(args.board, 'stub', args.stub_mode, games, 0, threads)
        if step >= args.warmup:
            tot_sims += r['sims']
            tot_plies += r['plies']
            tot_secs += r['seconds']
            sample = r['sample']
        if args.evaluator == 'net' and step >= 1 and tot_secs > 240:
            break
    value = tot_sims / tot_secs
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * tot_secs / max(1, args.steps),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args), 'board_size': args.board,
                   **SEARCH},
        'moves_per_sec': tot_plies / tot_secs,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads,
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print_json(line)


# ------------------------------------------------------------------ ours ----

def select_bytes(d, n):
    """Algorithmic bytes of the select kernel from device counters, with the
    reference's field widths: per level 12 B per child (N, W, P)
    + 8 B node link; virtual loss apply + undo 2 x (8 R + 8 W) per level;
    root masks 2 x ceil(n^2/32) x 4 B per descent; 4 B path entry per level."""
    nw = (n * n + 31) // 32
    return (12 * d['sum_children'] + 8 * d['sum_depth'] + 32 * d['sum_depth']
            + 8 * nw * d['simulations'] + 4 * d['sum_depth'])


def run_ours(args):
    import torch
    import torch.distributed as dist
    from azalea_b200 import LockstepSelfPlay, StubEvaluator
    from azalea_b200.network import HexNetwork

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a GPU (no CPU fallback); '
                           'use --impl reference for the CPU baseline')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured' if 'hbm_gbs' in peaks else 'fallback'

    if args.evaluator == 'net':
        torch.manual_seed(0)
        evaluator = HexNetwork(args.board, 6, 64).eval().to(dev)
        evaluator.prepare_inference(torch.bfloat16)
    else:
        evaluator = StubEvaluator(args.stub_mode)
    sp = LockstepSelfPlay(evaluator, num_games=args.games, board_size=args.board,
                          seed=0xBAD5EED5, rank=rank, world_size=world,
                          device=dev, cuda_graph=not args.no_graph,
                          collect_replay=True,
                          streams=args.streams if args.evaluator == 'net' else 1,
                          nodes_per_game=args.nodes_per_game or None, **SEARCH)
    G, per_move = args.games, sp.sims_per_move
    n_streams = sp.streams

    # ---- warm-up (also captures the CUDA graph) ----
    for _ in range(max(3, args.warmup)):
        sp.step_move()
    barrier()

    # ---- timed region 1: device-resident, K steps ----
    c0 = sp.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            sp.step_move()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    c1 = sp.counters()
    d_run = {k: c1[k] - c0[k] for k in c1}
    total_sims = world * G * per_move * args.steps
    value = total_sims / (ms * 1e-3)
    moves_per_sec = world * G * args.steps / (ms * 1e-3)

    # ---- timed region 2: end to end through the host-facing call ----
    # per step: new evaluator weights arrive from pinned host memory (the
    # trainer's hand-off; the reference pickles the whole agent per game,
    # parallel_player.py:36-38), the move runs, and the step's results --
    # every game's (move, move_id, result, ply) and the replay rows of the
    # games that finished -- are read back to the host.
    h2d = d2h = 0
    if args.evaluator == 'net':
        host_w = [p.detach().float().cpu().pin_memory() for p in evaluator.parameters()]
        h2d_step = sum(w.numel() * 4 for w in host_w)
    else:
        host_w, h2d_step = [], 0
    chosen_host = torch.zeros(G, 4, dtype=torch.int32).pin_memory()
    sp.harvest(gather=world > 1)        # untimed: first use sets up the exchange
    barrier()
    t0 = time.perf_counter()
    rows_out = 0
    for _ in range(args.steps):
        if host_w:
            with torch.no_grad():
                for p, w in zip(evaluator.parameters(), host_w):
                    p.copy_(w, non_blocking=True)
            evaluator.prepare_inference(torch.bfloat16)
            h2d += h2d_step
        sp.step_move()
        chosen_host.copy_(sp.chosen, non_blocking=True)
        rows = sp.harvest(gather=world > 1)  # syncs; D2H of finished games' rows (rank 0 gets all)
        rows_out += len(rows)
        d2h += chosen_host.numel() * 4 + rows.nbytes + 8
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_sims / e2e_s

    # ---- kernel leg: per-launch timing of our kernels with CUDA events ----
    eng = sp.eng
    stream = torch.cuda.current_stream()
    sel_ev, exp_ev = [], []
    cs0 = sp.counters()
    sel_ms = exp_ms = 0.0
    conv_ev = None
    if args.evaluator == 'net' and getattr(evaluator, 'tower', None) == 'tcgen05':
        conv_ev = evaluator.conv_events = []
    eng.select_root()
    sp._evaluate(True)
    eng.expand_root(None, 1 if args.evaluator == 'net' else 0)
    for _ in range(sp.num_batches):
        a, b, c, d = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        a.record(stream)
        eng.select(sp.batch, sp.coef, sp.noise_scale, sp.noise_alpha)
        b.record(stream)
        v, p, kind = sp._evaluate(False)
        c.record(stream)
        eng.expand_backup(v, p, kind)
        d.record(stream)
        sel_ev.append((a, b))
        exp_ev.append((c, d))
    eng.play_commit(sp.temperature, sp.depth, sp.move_sampling, True, True, sp.chosen)
    torch.cuda.synchronize()
    if conv_ev is not None:
        evaluator.conv_events = None
    sel_ms = sum(a.elapsed_time(b) for a, b in sel_ev)
    exp_ms = sum(a.elapsed_time(b) for a, b in exp_ev)
    cs1 = sp.counters()
    dk = {k: cs1[k] - cs0[k] for k in cs1}
    sel_bytes = select_bytes(dk, args.board)
    launches = sp.num_batches
    sel_avg_ms = sel_ms / launches
    achieved = sel_bytes / launches / (sel_avg_ms * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
        key = f'{args.board}x{args.board}/{G}/' + ('noise' if sp.noise_scale else 'nonoise')
        traffic = tj['k_select'][key]['traffic_bytes']
    except Exception:
        pass
    roofline = {
        'kernel': 'k_select', 'bound': 'hbm', 'achieved': achieved,
        'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak,
        'traffic': traffic, 'peak_source': peak_src,
        'avg_launch_ms': sel_avg_ms,
        'alg_bytes_per_launch': sel_bytes / launches,
        'alg_bytes_per_sim': sel_bytes / max(1, dk['simulations']),
        'mean_depth': dk['sum_depth'] / max(1, dk['simulations']),
        'mean_children': dk['sum_children'] / max(1, dk['sum_depth']),
        'select_share_of_step': sel_ms / (ms / args.steps),
        'expand_backup_avg_launch_ms': exp_ms / launches,
        'note': 'latency/occupancy-bound pointer chasing; see DESIGN.md',
    }
    roofline_tree = None
    if conv_ev:
        # the dominant kernel of the step with the network evaluator: k_conv3x3
        # (csrc/az_tower.cuh).  Algorithmic bytes per board and launch:
        # n*n cells x 64 channels x 2 B read + as much written, + as much again
        # for the residual of the second convolution of a block.
        cell_bytes = args.board * args.board * 128
        cbytes = sum(boards * cell_bytes * (3 if res else 2) for _, _, res, boards in conv_ev)
        cflop = sum(boards * args.board * args.board * 2 * 64 * 576 for _, _, _, boards in conv_ev)
        cms = sum(a.elapsed_time(b) for a, b, _, _ in conv_ev)
        conv_traffic = None
        try:
            conv_traffic = tj['k_conv3x3'][f'{args.board}x{args.board}/{G * sp.batch}']['traffic_bytes']
        except Exception:
            pass
        roofline_tree = roofline
        achieved_c = cbytes / (cms * 1e-3) / 1e9
        roofline = {
            'kernel': 'k_conv3x3', 'bound': 'hbm', 'achieved': achieved_c,
            'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved_c / hbm_peak,
            'traffic': conv_traffic, 'peak_source': peak_src,
            'launches_timed': len(conv_ev), 'avg_launch_ms': cms / len(conv_ev),
            'alg_bytes_per_launch': cbytes / len(conv_ev),
            'avg_launch_ms_plain': (sum(a.elapsed_time(b) for a, b, r, _ in conv_ev if not r)
                                    / max(1, sum(1 for e in conv_ev if not e[2]))),
            'avg_launch_ms_residual': (sum(a.elapsed_time(b) for a, b, r, _ in conv_ev if r)
                                       / max(1, sum(1 for e in conv_ev if e[2]))),
            'useful_tflops': cflop / (cms * 1e-3) / 1e12,
            'tensor_peak_tflops_sustained': float(peaks.get('bf16_tflops_sustained', 1400.0)),
            'conv_share_of_step': cms / (ms / args.steps),
            'note': 'two launches per residual block: x->y (2 units of traffic) and y,x->x (3 units); '
                    'the second sits on the HBM roofline, the first between HBM and the tensor pipe',
        }
    nn_info = None
    if args.evaluator == 'net':
        step_ms = ms / args.steps
        rows = d_run['nn_rows'] / args.steps
        padded = G * (per_move + 1)
        flop = NN_FLOP_PER_LEAF.get(args.board)
        bf16_peak = float(peaks.get('bf16_tflops_sustained', 1400.0))
        if flop:
            nn_ms = step_ms - (sel_ms + exp_ms)
            nn_info = {'useful_rows_per_step': rows, 'padded_rows_per_step': padded,
                       'tflops_padded': padded * flop / (nn_ms * 1e-3) / 1e12,
                       'tflops_useful': rows * flop / (nn_ms * 1e-3) / 1e12,
                       'peak_bf16_tflops_sustained': bf16_peak,
                       'approx_nn_ms_per_step': nn_ms}

    # ---- tree-only figure (stub evaluator) next to the headline ----
    tree_only = None
    if args.evaluator == 'net' and not args.skip_tree_only:
        del sp
        torch.cuda.empty_cache()
        sp2 = LockstepSelfPlay(StubEvaluator(2), num_games=args.games,
                               board_size=args.board, seed=1, rank=rank,
                               world_size=world, device=dev, cuda_graph=True,
                               collect_replay=True,
                               nodes_per_game=args.nodes_per_game or None, **SEARCH)
        for _ in range(3):
            sp2.step_move()
        barrier()
        k2 = max(args.steps, 8)
        ev0.record()
        for _ in range(k2):
            sp2.step_move()
        ev1.record()
        barrier()
        ms2 = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms2], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        tree_only = {'value': world * G * per_move * k2 / (ms2 * 1e-3), 'unit': UNIT,
                     'evaluator': 'device stub (mode 2)', 'ms_per_step': ms2 / k2,
                     'moves_per_sec': world * G * k2 / (ms2 * 1e-3)}
        del sp2

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        # bounded sample, ~10-30 s of CPU work: one ply of 4 games per core with
        # the network, 16 whole games per core with the stub evaluator
        games = args.cpu_games or (4 * cores if args.evaluator == 'net' else 16 * cores)
        r = cpu_selfplay_sample(args.board, args.evaluator, args.stub_mode, games,
                                1, cores)
        cpu_baseline = {'value': r['sims'] / r['seconds'], 'unit': UNIT,
                        'cores': cores, 'kind': 'port', 'sample': r['sample'],
                        'seconds': r['seconds']}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 tree statistics (bit-exact PUCT); bf16 network'
            if args.evaluator == 'net' else 'f32',
            'data': 'synthetic',
            'config': {'workload': workload_name(args), 'board_size': args.board,
                       'games_per_gpu': G, 'sims_per_move': per_move,
                       'l2_policy': 'working set (node pools, > 1 GB) exceeds the 126 MB L2',
                       'cuda_graph': not args.no_graph, 'streams': n_streams, **SEARCH},
            'moves_per_sec': moves_per_sec,
            'clocks': clk.summary(),
            'e2e': {'value': e2e_value, 'unit': UNIT,
                    'h2d_bytes_per_step': h2d // max(1, args.steps),
                    'd2h_bytes_per_step': d2h // max(1, args.steps),
                    'ms_per_step': 1e3 * e2e_s / args.steps,
                    'replay_rows_per_step': rows_out / args.steps},
            'gpu_launches': args.steps * sp_launches(args, per_move, evaluator),
            'roofline': roofline,
            'roofline_tree': roofline_tree,
            'cpu_baseline': cpu_baseline,
            'tree_only': tree_only,
            'network': nn_info,
            'counters_per_step': {k: v / args.steps for k, v in d_run.items()},
        }
        print_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sp_launches(args, per_move, evaluator=None):
    """Kernels of ours per step: select_root, expand_root, commit, and per
    search batch select + expand_backup, plus per evaluation (root + every
    search batch) the stub kernel, or the network's stem + heads kernels and,
    with the tcgen05 tower, two k_conv3x3 per residual block."""
    nb = per_move // SEARCH['search_batch_size']
    per = 3 + 2 * nb
    if args.evaluator == 'stub':
        per += nb + 1
    else:
        own = 2
        if getattr(evaluator, 'tower', None) == 'tcgen05':
            own += 2 * len(evaluator.resblocks)
        per += (nb + 1) * own
    return per


def main():
    args = parse_args()
    # libraries (NCCL's version banner) may write to stdout: keep fd 1 for
    # the ONE JSON line and send everything else to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global print_json

    def print_json(line):
        os.write(json_fd, (json.dumps(line) + '\n').encode())

    if args.impl == 'reference':
        rank = int(os.environ.get('RANK', 0))
        if rank != 0:
            return
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
