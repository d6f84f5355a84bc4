"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: row-block partition, the single
B broadcast, and that concatenated row-blocks equal the single-rank result bit for bit.  The
compute callback here is the ORACLE (test infrastructure) standing in for the single-GPU launch."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gemm_hls_b200 import multi  # noqa: E402


def test_row_block_partition_covers_all_rows_once():
    for n in (1, 7, 8, 513, 8192, 16384):
        for world in (1, 2, 3, 4, 8):
            blocks = [multi.row_block(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 <= a1
            assert sum(b[1] - b[0] for b in blocks) == n
    assert multi.row_block(8192, 8, 3) == (3072, 4096)   # BASELINE config 4: 1024 rows per GPU
    with pytest.raises(ValueError):
        multi.row_block(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, k, m, mp, rd, out_dir):
    import torch
    import torch.distributed as dist
    import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = O.fill(O.FLOAT, n, k, m)
    r0, r1 = multi.row_block(n, world, rank)
    a_blk = torch.from_numpy(a.reshape(n, k)[r0:r1].copy())
    # only rank 0 holds B before the broadcast
    b_t = torch.from_numpy(b.reshape(k, m).copy()) if rank == 0 else torch.zeros((k, m), dtype=torch.float32)

    def compute(a_block, b_full):
        rows = a_block.shape[0]
        if rows == 0:
            return torch.zeros((0, m), dtype=torch.float32)
        return torch.from_numpy(O.naive(O.FLOAT, mp, rd, a_block.numpy(), b_full.numpy(), rows, k, m))

    c_blk = multi.rowblock_matmul(a_blk, b_t, compute)
    assert torch.equal(b_t, torch.from_numpy(b.reshape(k, m)))      # B arrived everywhere
    full = multi.gather_rows(c_blk, n)
    np.save(os.path.join(out_dir, "c_rank%d.npy" % rank), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,k,m,mp,rd", [(37, 32, 48, 0, 1), (64, 64, 32, 1, 2)])
def test_two_rank_row_block_split_equals_single_rank(tmp_path, n, k, m, mp, rd):
    import torch.multiprocessing as tmp_mp
    import oracle as O
    port = _free_port()
    tmp_mp.spawn(_worker, args=(2, port, n, k, m, mp, rd, str(tmp_path)), nprocs=2, join=True)
    a, b = O.fill(O.FLOAT, n, k, m)
    single = O.naive(O.FLOAT, mp, rd, a, b, n, k, m)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "c_rank%d.npy" % r))
        assert got.tobytes() == single.tobytes()
