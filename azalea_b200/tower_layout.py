"""The tcgen05 tower's activation / weight layouts in plain torch.

Reference implementation of what ``az_nn_stem`` writes, ``az_nn_conv3x3``
reads and writes and ``az_nn_heads`` reads (csrc/az_tower.cuh).  Used by the
tests and by ``HexNetwork.prepare_inference`` for the weights; the hot path
never converts activations on the host.
"""
import torch

HALO = 8


def boards_per_group(n):
    return 128 // (n + 1)


def buffer_rows(n, num_boards):
    bpg = boards_per_group(n)
    groups = (num_boards + bpg - 1) // bpg
    return HALO + groups * n * 128 + 16


def swizzle_rows(t):
    """[R, 64] -> 16-byte chunk j of row R stored at chunk j ^ (R & 7)."""
    R = t.shape[0]
    v = t.reshape(R, 8, 8)
    idx = torch.arange(8, device=t.device)[None, :] ^ (torch.arange(R, device=t.device)[:, None] & 7)
    out = torch.empty_like(v)
    out.scatter_(1, idx[:, :, None].expand(R, 8, 8), v)
    return out.reshape(R, 64)


def unswizzle_rows(t):
    R = t.shape[0]
    idx = torch.arange(8, device=t.device)[None, :] ^ (torch.arange(R, device=t.device)[:, None] & 7)
    return torch.gather(t.reshape(R, 8, 8), 1, idx[:, :, None].expand(R, 8, 8)).reshape(R, 64)


def row_index(n, num_boards, device):
    """Global row of every real cell: int64 [num_boards, n, n]."""
    bpg = boards_per_group(n)
    b = torch.arange(num_boards, device=device)[:, None, None]
    y = torch.arange(n, device=device)[None, :, None]
    x = torch.arange(n, device=device)[None, None, :]
    return HALO + ((b // bpg) * n + y) * 128 + (b % bpg) * (n + 1) + x


def to_slabs(x):
    """[N, n, n, 64] (NHWC) -> slab-layout buffer [rows, 64], pre-swizzled."""
    N, n = x.shape[0], x.shape[1]
    buf = torch.zeros(buffer_rows(n, N), 64, dtype=x.dtype, device=x.device)
    buf[row_index(n, N, x.device).reshape(-1)] = x.reshape(-1, 64)
    return swizzle_rows(buf)


def from_slabs(buf, n, num_boards):
    """Slab-layout buffer -> ([N, n, n, 64] real cells, max |v| over every row
    that is not a board cell; the unused board slots of the last group count
    as cells -- the tower computes them like any board)."""
    t = unswizzle_rows(buf)
    idx = row_index(n, num_boards, buf.device).reshape(-1)
    real = t[idx].reshape(num_boards, n, n, 64)
    bpg = boards_per_group(n)
    slots = (num_boards + bpg - 1) // bpg * bpg
    mask = torch.ones(t.shape[0], dtype=torch.bool, device=buf.device)
    mask[row_index(n, slots, buf.device).reshape(-1)] = False
    rest = t[mask]
    return real, (float(rest.float().abs().max()) if rest.numel() else 0.0)


def pack_conv_weights(w):
    """[c_out, c_in, ky, kx] (64 x 64 x 3 x 3) -> [3 kx][3 ky][64 c_out][64 c_in],
    rows pre-swizzled by (row & 7) == (c_out & 7)."""
    t = w.permute(3, 2, 0, 1).reshape(9 * 64, 64).contiguous()
    return swizzle_rows(t)
