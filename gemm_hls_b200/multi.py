"""Row-block partition of the hot path across the GPUs of one box.

Outer tiles of C are independent in the reference (kernel/Compute.cpp:53-56: no cross-tile state,
C written once, kernel/Memory.cpp:361-392), so C and A are split into contiguous row-blocks, one
per rank; B is replicated with ONE broadcast from rank 0 before any compute (NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests); there is no per-step collective and no reduction
(K is not split).  One process per GPU, torch.distributed for the plumbing.
"""
from typing import Callable, Tuple


def row_block(size_n: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [r0, r1) of C (and A) owned by `rank`: ceil(N / world) rows each, the tail ranks may
    get fewer (or none when world > N)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = (size_n + world - 1) // world
    r0 = min(size_n, rank * per)
    return r0, min(size_n, r0 + per)


def broadcast_b(b, src: int = 0, group=None):
    """The path's one collective: replicate B (K x M) from `src` to every rank, in place."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(b, src=src, group=group)
    return b


def rowblock_matmul(a_block, b, compute: Callable, group=None):
    """C_block = compute(a_block, b) on this rank's row-block after B has been broadcast.
    `compute(a_block, b) -> c_block` is the single-GPU launch (the C-ABI call in production)."""
    broadcast_b(b, 0, group)
    return compute(a_block, b)


def gather_rows(c_block, size_n: int, group=None):
    """Concatenate the row-blocks on every rank (verification / host-side consumers only; the
    benchmark leaves C distributed)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return c_block
    world = dist.get_world_size(group)
    per = (size_n + world - 1) // world
    pad = torch.zeros((per,) + tuple(c_block.shape[1:]), dtype=c_block.dtype, device=c_block.device)
    pad[: c_block.shape[0]] = c_block
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat(parts, dim=0)[:size_n]
