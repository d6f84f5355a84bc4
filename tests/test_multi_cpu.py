"""CPU tests (gloo, world_size 2 and 4) of the multi-GPU host logic bench.py runs under torchrun: the r x c grid of
C blocks, the single B broadcast, and that the assembled blocks equal the single-rank result bit for bit.  The
compute callback here is the ORACLE (test infrastructure) standing in for the single-GPU launch."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gemm_hls_b200 import multi  # noqa: E402


def test_rank_grid_minimises_replicated_preparation():
    assert multi.rank_grid(1, 16384, 16384, 16384) == (1, 1)
    assert multi.rank_grid(2, 16384, 16384, 16384) == (2, 1)     # tie -> rows: the plain row-block split
    assert multi.rank_grid(4, 16384, 16384, 16384) == (2, 2)
    assert multi.rank_grid(8, 16384, 16384, 16384) == (4, 2)     # tie between (4, 2) and (2, 4) -> more rows
    assert multi.rank_grid(8, 8192, 8192, 8192) == (4, 2)
    assert multi.rank_grid(8, 64, 4096, 65536) == (1, 8)         # a wide B: split columns only
    assert multi.rank_grid(8, 65536, 4096, 64) == (8, 1)         # a tall A: rows only
    with pytest.raises(ValueError):
        multi.rank_grid(0, 1, 1, 1)


def test_blocks_cover_c_exactly_once():
    for n, m, width in ((1, 16, 16), (7, 48, 16), (513, 528, 16), (8192, 8192, 8), (16384, 16384, 16), (100, 96, 32)):
        for world in (1, 2, 3, 4, 6, 8):
            for r in range(1, world + 1):
                if world % r:
                    continue
                grid = (r, world // r)
                cover = np.zeros((n, m), dtype=np.int32)
                for rank in range(world):
                    r0, r1, c0, c1 = multi.rank_block(rank, grid, n, m, width)
                    assert 0 <= r0 <= r1 <= n and 0 <= c0 <= c1 <= m
                    assert (c1 - c0) % width == 0 and c0 % width == 0
                    cover[r0:r1, c0:c1] += 1
                assert np.all(cover == 1), (n, m, width, grid)
    assert multi.row_block(8192, 8, 3) == (3072, 4096)           # BASELINE config 4 as a row split: 1024 rows per GPU
    assert multi.rank_block(5, (4, 2), 16384, 16384, 16) == (8192, 12288, 8192, 16384)
    with pytest.raises(ValueError):
        multi.row_block(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, grid, n, k, m, mp, rd, out_dir):
    import torch
    import torch.distributed as dist
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = O.fill(O.FLOAT, n, k, m)
    block = multi.rank_block(rank, grid, n, m, 16)
    r0, r1, c0, c1 = block
    a_blk = torch.from_numpy(a.reshape(n, k)[r0:r1].copy())
    # only rank 0 holds B before the broadcast
    b_t = torch.from_numpy(b.reshape(k, m).copy()) if rank == 0 else torch.zeros((k, m), dtype=torch.float32)

    def compute(a_block, b_block):
        rows, cols = a_block.shape[0], b_block.shape[1]
        if rows == 0 or cols == 0:
            return torch.zeros((rows, cols), dtype=torch.float32)
        return torch.from_numpy(O.naive(O.FLOAT, mp, rd, a_block.numpy(), b_block.numpy(), rows, k, cols))

    c_blk = multi.block_matmul(a_blk, b_t, block, compute)
    assert torch.equal(b_t, torch.from_numpy(b.reshape(k, m)))      # B arrived everywhere
    assert tuple(c_blk.shape) == (r1 - r0, c1 - c0)
    full = multi.gather_blocks(c_blk, grid, n, m, 16)
    np.save(os.path.join(out_dir, "c_rank%d.npy" % rank), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,grid,n,k,m,mp,rd", [(2, (2, 1), 37, 32, 48, 0, 1), (2, (1, 2), 64, 64, 32, 1, 2),
                                                    (4, (2, 2), 37, 32, 48, 0, 1), (4, (4, 1), 3, 16, 16, 0, 1)])
def test_block_split_equals_single_rank(tmp_path, world, grid, n, k, m, mp, rd):
    import torch.multiprocessing as tmp_mp
    import oracle as O
    port = _free_port()
    tmp_mp.spawn(_worker, args=(world, port, grid, n, k, m, mp, rd, str(tmp_path)), nprocs=world, join=True)
    a, b = O.fill(O.FLOAT, n, k, m)
    single = O.naive(O.FLOAT, mp, rd, a, b, n, k, m)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "c_rank%d.npy" % r))
        assert got.tobytes() == single.tobytes()
