"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame sharding and the match key reduction."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import yolopoint_oracle as O
from yolopoint_b200.dist import reduce_match_keys, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _keys_numpy(d1, d2, col_off):
    """numpy statement of what yp_match_partial emits (include/yolopoint_b200.h): packed (dist bits << 32 | index) minima."""
    dm = np.sqrt(2 - 2 * np.clip(d1.T @ d2, -1, 1)).astype(np.float32)
    bits = dm.view(np.uint32).astype(np.int64) << 32
    row = (bits | (np.arange(d2.shape[1], dtype=np.int64)[None, :] + col_off)).min(1)
    col = (bits | np.arange(d1.shape[1], dtype=np.int64)[:, None]).min(0)
    return row, col


def _finalize_numpy(row, col, thr):
    j = row & 0xffffffff
    dist_ = (row >> 32).astype(np.uint32).view(np.float32)
    keep = (dist_ < thr) & ((col[j] & 0xffffffff) == np.arange(len(row)))
    return np.stack((np.arange(len(row))[keep], j[keep], dist_[keep])).astype(np.float64)


def _worker(rank, world, port, d1, d2, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(d2.shape[1], rank, world)
    row, col = _keys_numpy(d1, d2[:, lo:hi], lo)
    rk, ck = reduce_match_keys(torch.from_numpy(row), torch.from_numpy(col), d2.shape[1])
    q.put((rank, _finalize_numpy(rk.numpy(), ck.numpy(), 0.7)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_match_reduction_gloo():
    rs = np.random.RandomState(0)
    D, N1, N2 = 32, 203, 157
    d1 = rs.normal(0, 1, (D, N1)).astype(np.float32); d1 /= np.linalg.norm(d1, axis=0)
    d2 = rs.normal(0, 1, (D, N2)).astype(np.float32)
    d2[:, :100] = d1[:, rs.permutation(N1)[:100]] + 0.05 * rs.normal(0, 1, (D, 100)).astype(np.float32)
    d2 /= np.linalg.norm(d2, axis=0)
    ref = O.nn_match_two_way(d1, d2, 0.7)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, d1, d2, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for _, m in outs:
        np.testing.assert_array_equal(m[:2], ref[:2])
        np.testing.assert_allclose(m[2], ref[2], rtol=0, atol=1e-6)
