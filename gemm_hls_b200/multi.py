"""Partition of the hot path across the GPUs of one box for the one-process-per-GPU launch (bench.py under
torchrun).  The in-library single-process split is mm_multi_* (csrc/capi.cu).

Outer tiles (n0, m0) of C are independent in the reference (kernel/Compute.cpp:53-56: no cross-tile state, C
written once, kernel/Memory.cpp:361-392), so C is cut over an r x c grid of ranks: rank (i, j) computes row-block
i x column-block j of C from A's row-block i and B's column-block j.  B is distributed with ONE broadcast from rank
0 before any compute (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests); there is no per-step collective and
no reduction (K is not split).  c = 1 is the plain row-block split with B replicated; a 2-D grid replicates less
operand preparation per step (a rank prepares 1/r of A and 1/c of B).
"""
from typing import Callable, Tuple


def rank_grid(world: int, n: int, k: int, m: int) -> Tuple[int, int]:
    """(r, c) with r * c == world minimising the operand elements a rank prepares per step, n*k/r + k*m/c;
    ties go to more row-blocks, so world = 2 on a square problem stays the row-block split."""
    if world <= 0:
        raise ValueError("bad world size")
    best = None
    for r in range(1, world + 1):
        if world % r:
            continue
        c = world // r
        cost = n * k / r + k * m / c
        if best is None or cost < best[0] - 1e-9 or (abs(cost - best[0]) <= 1e-9 and r > best[1]):
            best = (cost, r, c)
    return best[1], best[2]


def rank_block(rank: int, grid: Tuple[int, int], n: int, m: int, width: int = 1) -> Tuple[int, int, int, int]:
    """(r0, r1, c0, c1): rows [r0, r1) and columns [c0, c1) of C owned by `rank` in an r x c grid (row-major rank
    order).  ceil-sized blocks, tail blocks may be short or empty; column blocks are multiples of `width` elements
    (the 64-byte memory word of the data type: a block's M must satisfy the reference's shape rule too)."""
    r, c = grid
    if not (0 <= rank < r * c):
        raise ValueError("bad rank for this grid")
    i, j = rank // c, rank % c
    rows_per = (n + r - 1) // r
    cols_per = ((m + c - 1) // c + width - 1) // width * width
    r0, c0 = min(n, i * rows_per), min(m, j * cols_per)
    return r0, min(n, r0 + rows_per), c0, min(m, c0 + cols_per)


def row_block(size_n: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [r0, r1) of the plain row-block split (the c = 1 grid)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    r0, r1, _, _ = rank_block(rank, (world, 1), size_n, 1)
    return r0, r1


def broadcast_b(b, src: int = 0, group=None):
    """The path's one collective: replicate B (K x M) from `src` to every rank, in place."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(b, src=src, group=group)
    return b


def local_b(b, c0: int, c1: int):
    """This rank's column-block of B as its own dense K x (c1 - c0) array (the kernels take no leading dimension,
    like the reference); the whole of B when the block is all of it."""
    if c0 == 0 and c1 == b.shape[1]:
        return b
    return b[:, c0:c1].contiguous()


def block_matmul(a_block, b, block, compute: Callable, group=None):
    """C_block = compute(a_block, B[:, c0:c1]) after B has been broadcast; `compute(a_block, b_block) -> c_block`
    is the single-GPU launch (the C-ABI call in production)."""
    broadcast_b(b, 0, group)
    return compute(a_block, local_b(b, block[2], block[3]))


def gather_blocks(c_block, grid: Tuple[int, int], n: int, m: int, width: int = 1, group=None):
    """Assemble the full C on every rank (verification / host-side consumers only; the benchmark leaves C
    distributed)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return c_block
    world = dist.get_world_size(group)
    r, c = grid
    rows_per = (n + r - 1) // r
    cols_per = ((m + c - 1) // c + width - 1) // width * width
    pad = torch.zeros((rows_per, cols_per), dtype=c_block.dtype, device=c_block.device)
    pad[: c_block.shape[0], : c_block.shape[1]] = c_block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    full = torch.empty((n, m), dtype=c_block.dtype, device=c_block.device)
    for rank, part in enumerate(parts):
        r0, r1, c0, c1 = rank_block(rank, grid, n, m, width)
        full[r0:r1, c0:c1] = part[: r1 - r0, : c1 - c0]
    return full
