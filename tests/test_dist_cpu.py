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


# ---------------------------------------------------------------------------------------------------------------------
# training step: flat-buffer gradient all-reduce (training row)
# ---------------------------------------------------------------------------------------------------------------------
def _train_worker(rank, world, port, q):
    import torch
    from yolopoint_b200 import Model
    from yolopoint_b200.trainer import TrainStep, synthetic_sample
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    torch.manual_seed(0)
    m = Model(names=[str(i) for i in range(80)], version="n")            # identical weights on every rank
    sample = synthetic_sample(2, 64, 96, seed=10 + rank)                  # a different batch slice per rank
    ts = TrainStep(m, sparse_cfg=dict(num_samples_per_image=40, num_masked_non_matches_per_match=10), bucket_bytes=1 << 20)
    assert len(ts.reducer.buckets) > 3
    state = {k: v.clone() for k, v in m.state_dict().items()}
    # expectation: mean over ranks of the purely local gradients, computed with plain autograd on a fresh model copy
    torch.manual_seed(0)
    m2 = Model(names=[str(i) for i in range(80)], version="n")
    m2.load_state_dict(state)
    m2.train()
    from yolopoint_b200 import losses as Lz
    from yolopoint_b200.trainer import LAMBDA_DESC, LAMBDA_OBJ, LOSS_CFG
    ref_step = TrainStep.__new__(TrainStep)
    ref_step.model, ref_step.device = m2, torch.device("cpu")
    ref_step.obj_loss, ref_step.det_loss = Lz.ComputeObjectLoss(m2, LOSS_CFG, "cpu"), Lz.ComputeDetectorLoss("cpu")
    ref_step.sparse_cfg, ref_step.graphed, ref_step.desc_loss = ts.sparse_cfg, None, ts.desc_loss
    torch.manual_seed(100 + rank)
    l2, _ = ref_step.losses(sample)
    l2.backward()
    g_local = torch.cat([p.grad.flatten() for p in m2.parameters()])
    both = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(both, g_local)
    expect = sum(both) / world
    # the real step
    m.load_state_dict(state)
    torch.manual_seed(100 + rank)
    before = [p.detach().clone() for p in m.parameters()]
    loss = ts.step(sample)
    got = ts.reducer.flat.clone()
    changed = sum(int(not torch.equal(a, p.detach())) for a, p in zip(before, m.parameters()))
    q.put((rank, float((got - expect).abs().max()), float(expect.abs().max()), float(loss), changed, len(before), len(ts.reducer.handles)))
    dist.barrier()
    dist.destroy_process_group()


def test_training_step_gradient_allreduce_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for rank, err, scale, loss, changed, n_params, n_handles in outs:
        assert np.isfinite(loss) and n_handles > 3                    # one asynchronous all-reduce per bucket
        assert err <= 1e-5 * max(scale, 1.0), (rank, err, scale)     # averaged gradients == mean of the per-rank gradients
        assert changed > 0.9 * n_params                                # Adam moved (nearly) every parameter
