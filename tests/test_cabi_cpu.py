"""CPU-side checks of the C-ABI library: it loads without a GPU and exports
every symbol include/azalea_b200.h declares.  No compute calls."""
import ctypes as C

import pytest

from azalea_b200 import _cabi


def test_library_loads_and_exports_all_declared_symbols():
    L = _cabi.lib()
    names = _cabi.declared_symbols()
    assert len(names) >= 24
    for name in names:
        assert hasattr(L, name), name
    assert L.az_abi_version() == 3


def test_device_bytes_is_pure_host_arithmetic():
    L = _cabi.lib()
    cfg = _cabi.AzConfig(num_games=4096, board_size=11, max_batch=10,
                         nodes_per_game=200_000, max_nodes_ref=10_000_000,
                         replay_rows=100_000, max_plies=300, seed=1,
                         first_game_id=0, game_id_stride=4096)
    nbytes = L.az_engine_device_bytes(C.byref(cfg))
    # dominated by the ping-pong node pool: G * 2 * C * 16 bytes
    assert nbytes > 4096 * 2 * 200_000 * 16
    assert nbytes < 4096 * 2 * 200_000 * 16 * 1.1
    for bad in (dict(board_size=20), dict(board_size=1), dict(max_batch=33),
                dict(num_games=0), dict(nodes_per_game=1 << 24)):
        kw = dict(num_games=8, board_size=11, max_batch=10,
                  nodes_per_game=1000, max_nodes_ref=0, replay_rows=0,
                  max_plies=0, seed=0, first_game_id=0, game_id_stride=0)
        kw.update(bad)
        assert L.az_engine_device_bytes(C.byref(_cabi.AzConfig(**kw))) == 0


def test_error_strings():
    L = _cabi.lib()
    assert L.az_strerror(0) == b'ok'
    assert b'invalid' in L.az_strerror(-1)
    with pytest.raises(RuntimeError):
        _cabi.check(-1)


def test_no_cpu_fallback_without_cuda():
    """On a box without a GPU the product refuses to run instead of
    falling back to a CPU implementation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from azalea_b200 import HexGame, Engine
    with pytest.raises(RuntimeError):
        HexGame(11)
    with pytest.raises(RuntimeError):
        Engine(1, 11, device='cpu')
