"""BASELINE.json configs[3]: brute-force two-way descriptor match sweep, 512..16384 keypoints per frame, 1 vs N GPUs.

  python tools/bench_match.py                      # one GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_match.py

desc1 is replicated, desc2 is sharded by columns over the ranks; the exchange step is an integer MIN all-reduce of packed
(distance, index) keys plus an all-gather of the column keys (yolopoint_b200/dist.py).  Prints one JSON line per size."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from yolopoint_b200 import ops  # noqa: E402
from yolopoint_b200.dist import match_two_way_sharded  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    D = int(os.environ.get("YP_MATCH_D", "256"))
    for N in (512, 1024, 2048, 4096, 8192, 16384):
        g = torch.Generator(device="cpu").manual_seed(N)
        d1 = torch.randn(N, D, generator=g); d1 /= d1.norm(dim=1, keepdim=True)
        perm = torch.randperm(N, generator=g)
        d2 = d1[perm] + 0.05 * torch.randn(N, D, generator=g); d2 /= d2.norm(dim=1, keepdim=True)   # planted matches
        d1, d2 = d1.to(dev), d2.to(dev).contiguous()
        algo = os.environ.get("YP_MATCH_ALGO", "auto")
        if world > 1:
            fn = lambda: match_two_way_sharded(d1, d2, 0.7)
        elif os.environ.get("YP_MATCH_GRAPH", "1") != "0":
            matcher = ops.TwoWayMatcher(N, N, D, dev, 0.7, algo=algo)     # static operands + one CUDA graph per call
            matcher.d1.copy_(d1); matcher.d2.copy_(d2)
            fn = lambda m=matcher: m(m.d1, m.d2)
        else:
            fn = lambda: ops.match_two_way(d1, None, d2, None, 0.7, algo=algo)
        for _ in range(3):
            m, cnt = fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 20 if N <= 4096 else 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            m, cnt = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            t = float(ms) * 1e-3
            flops = 2.0 * N * N * D
            print(json.dumps({"workload": f"two-way match N1=N2={N} D={D}", "n_gpus": world, "ms": float(ms), "matches": int(cnt.item()),
                              "tflops_fp32": flops / t / 1e12, "compulsory_GBps": (2 * N * D * 4 + N * 16) / t / 1e9}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
