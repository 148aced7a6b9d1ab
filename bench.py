#!/usr/bin/env python
"""Benchmark of the YOLOPoint hot path on B200 (contract: see the task statement / DESIGN.md section "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload s640|n480|m1280|l640train|s640v52] [--precision fp32|bf16]

A "step" is one pass of the whole per-frame hot path (uint8 frame -> network -> Detect decode -> box NMS -> heatmap ->
keypoint NMS -> descriptor sampling -> two-way match with the previous frame) over one batch of synthetic frames.
Default workload = BASELINE.json configs[1]: YOLOPoint-S, 640x640, batch 1 per GPU.  N > 1 (torchrun) shards
independent frames over ranks with no data-path collective (weak scaling).

One JSON line is printed by rank 0.  `value` = frames/s with the input frames resident in HBM; `e2e` = frames/s through
FramePipeline.submit_host/collect (pinned host frame -> H2D -> pipeline -> D2H of keypoints/descriptors/boxes/matches; one camera
stream, the host runs one frame ahead);
`roofline` describes the dominant kernel (tcgen05 conv) against the measured bf16 tensor peak; `cpu_baseline` is the CPU
oracle (a port of the reference's PyTorch/numpy path) timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    #  name: (version, H, W, frames per GPU per step)
    "s640": ("s", 640, 640, 1),     # BASELINE.json configs[1]
    "n480": ("n", 480, 640, 1),     # configs[0] geometry (the reference's CPU case) on the GPU
    "m1280": ("m", 736, 1280, 4),   # configs[2]: 32 frames over 8 GPUs = 4 per GPU
    "l640train": ("l", 640, 640, 8),  # configs[4]: YOLOPoint-L bf16 training step, 64 samples over 8 GPUs = 8 per GPU
    "s640v52": ("s", 640, 640, 1),  # YOLOPointv52-S (the model configs/kitti_inference.yaml names), configs[1] geometry
}
MODEL_NAME = {"s640v52": "YOLOPointv52"}   # every other workload runs the YOLOPoint (v5-style) network
CONV_DRAM_BYTES_PER_LAUNCH = {"s640": 6.43e6}   # profiles/r02_drain_pass_dram.md: dram read + write per launch, mean over the 64 conv launches of one pass of this plan (cold L2)
NAMES = [str(i) for i in range(80)]
OTHER_CONFIGS_BUDGET_S = 300              # bound on the child runs of the other BASELINE configs reported under detail
E2E_REPEATS = 3                           # the end-to-end region is repeated and its median reported (see main())
# SURVEY.md section 8a, per frame (forward); s640v52: conv-module hook count on the reference YOLOPointv52-S
CONV_GFLOP = {"s640": 21.023, "n480": 4.232, "m1280": 141.398, "l640train": 135.526, "s640v52": 21.363}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops_sustained"], tf_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, src="fallback")


def build_weights(version, model_name="YOLOPoint"):
    from yolopoint_b200 import Model
    from yolopoint_b200.synth import perturb_state_dict
    torch.manual_seed(0)
    m = Model(names=NAMES, version=version, model_name=model_name)
    sd = perturb_state_dict(m.state_dict(), 0, version)
    m.load_state_dict(sd)
    return m, sd


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, n): n.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", "")
                 for n in dir(nv) if n.startswith("nvmlClocksThrottleReason") or n.startswith("nvmlClocksEventReason")}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if isinstance(bit, int) and bit and (r & bit) and n not in ("None", "All", "GpuIdle"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_fps(version, H, W, steps, warmup, sd=None, model_name="YOLOPoint"):
    """The reference's CPU path (oracle port: same torch-CPU convs / numpy post-processing).  The thread count is tuned first: the
    cores this process may use (sched_getaffinity, not cpu_count) are an upper bound, and on shared hosts fewer threads can be much
    faster (r01: 6.3 frames/s with 16 threads, 19 with torchrun's OMP_NUM_THREADS=1), so a few counts are tried on two frames each
    and the best one is used and reported.  -> (frames/s, threads used, median seconds per frame, {threads: frames/s})"""
    from oracle import yolopoint_oracle as O
    from yolopoint_b200.synth import synthetic_frame
    if sd is None:
        _, sd = build_weights(version, model_name)
    net = O.OracleNet(sd, version, 80, model_name)
    frames = [synthetic_frame(H, W, s) for s in range(2)]
    state = {"prev": None}

    def one(i):
        t0 = time.perf_counter()
        pts, desc, boxes = O.process_frame(net, frames[i % 2])
        if state["prev"] is not None and desc is not None:
            O.nn_match_two_way(state["prev"], desc, O.DEFAULT_CFG["nn_thresh"])
        state["prev"] = desc
        return time.perf_counter() - t0

    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sweep = {}
    for n in sorted({1, 2, 4, 8, 16, 32, avail} & set(range(1, avail + 1))):
        torch.set_num_threads(n)
        one(0)
        sweep[n] = 2.0 / (one(1) + one(0))
    best = max(sweep, key=sweep.get)
    torch.set_num_threads(best)
    times = []
    for i in range(warmup + steps):
        dt = one(i)
        if i >= warmup:
            times.append(dt)
    return len(times) / sum(times), best, float(np.median(times)), {str(k): round(v, 2) for k, v in sweep.items()}


def torch_eager_net_ms(sd, version, model_name, B, H, W, dev):
    """The network alone through eager PyTorch on the same GPU: this repo's module tree (the reference's layer-for-layer graph, BN
    folded like the reference's `.eval().fuse()`) on cuDNN / ATen kernels -- what the reference itself would run on a CUDA device
    (SURVEY.md section 2b).  Two settings: PyTorch's default (cuDNN may use plain TF32 for fp32 convolutions) and true fp32."""
    from yolopoint_b200 import Model
    tm = Model(names=NAMES, version=version, model_name=model_name)
    tm.load_state_dict(sd)
    tm = tm.to(dev).fuse().eval()
    x = torch.rand(B, 3, H, W, device=dev)
    out = {}
    prev = torch.backends.cudnn.allow_tf32
    try:
        for key, allow in (("ms_cudnn_default_tf32_allowed", True), ("ms_cudnn_fp32", False)):
            torch.backends.cudnn.allow_tf32 = allow
            with torch.no_grad():
                for _ in range(5):
                    tm.model(x)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    tm.model(x)
                e1.record()
                torch.cuda.synchronize(dev)
            out[key] = e0.elapsed_time(e1) / 20
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    out["what"] = "network forward only (no decode / NMS / keypoints / match), eager launches, batch as in config.workload"
    return out


def other_configs():
    """Short, bounded runs of the other BASELINE.json configs in child processes (their full JSON lines are kept under profiles/):
    configs[2] YOLOPoint-M 1280x736 batch 4 per GPU (bf16 and fp32), configs[3] descriptor match sweep, configs[4] YOLOPoint-L
    training step batch 8 per GPU."""
    import subprocess
    res = {}
    t_start = time.perf_counter()

    def child(tag, cmd, pick, timeout=420):
        left = OTHER_CONFIGS_BUDGET_S - (time.perf_counter() - t_start)
        if left <= 0:            # the default command stays within minutes whatever the box does: later children are skipped, not awaited
            res[tag] = {"skipped": f"time budget of {OTHER_CONFIGS_BUDGET_S} s for the other configs spent"}
            return
        timeout = min(timeout, left + 60)
        try:
            r = subprocess.run([sys.executable] + cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
            lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
            res[tag] = pick(lines) if lines else {"error": (r.stderr or "no output")[-300:]}
        except Exception as e:  # pragma: no cover
            res[tag] = {"error": str(e)[:300]}

    short = ["--steps", "20", "--warmup", "5", "--no-cpu-baseline", "--no-other-configs", "--also-streams", "0"]
    frames = lambda ls: {"frames_per_s": ls[-1]["value"], "e2e_frames_per_s": ls[-1]["e2e"]["value"], "ms_per_step": ls[-1]["ms_per_step"],
                         "net_only_ms": ls[-1]["detail"]["net_only_ms"], "conv_tflops": ls[-1]["roofline"]["achieved"],
                         "roofline_frac": ls[-1]["roofline"]["frac"], "workload": ls[-1]["config"]["workload"]}
    child("s640_one_frame_at_a_time", [os.path.join(ROOT, "bench.py"), "--in-flight", "1", "--tile-policy", "latency", "--steps", "100", "--warmup", "10",
                                       "--no-cpu-baseline", "--no-other-configs", "--also-streams", "0"],
          lambda ls: dict(frames(ls), what="the same workload with ONE frame in flight and the latency-tuned tile tables: ms_per_step is then the latency of a frame"))
    child("m1280_bf16", [os.path.join(ROOT, "bench.py"), "--workload", "m1280", "--precision", "bf16"] + short, frames)
    child("m1280_fp32", [os.path.join(ROOT, "bench.py"), "--workload", "m1280", "--precision", "fp32"] + short, frames)
    child("l640train_bf16", [os.path.join(ROOT, "bench.py"), "--workload", "l640train", "--steps", "5", "--warmup", "3"],
          lambda ls: {"samples_per_s": ls[-1]["value"], "ms_per_step": ls[-1]["ms_per_step"], "conv_tflops": ls[-1]["roofline"]["achieved"],
                      "roofline_frac": ls[-1]["roofline"]["frac"], "workload": ls[-1]["config"]["workload"]})
    child("match_sweep_d256", [os.path.join(ROOT, "tools", "bench_match.py")],
          lambda ls: [{"n": int(l["workload"].split("=")[2].split()[0]), "ms": l["ms"], "tflops_fp32_equiv": l["tflops_fp32"], "matches": l["matches"]} for l in ls])
    return res


def train_main(args, version, H, W, per_gpu, world, rank, local_rank):
    """configs[4]: one training step = 2 forwards + 3 losses + backward + gradient all-reduce + Adam (yolopoint_b200/trainer.py).
    value = samples/s over all ranks with the synthetic batch resident in HBM; e2e = the same step fed from pinned host memory
    (H2D of the sample inside the timed region, D2H of the loss)."""
    from yolopoint_b200 import Model
    from yolopoint_b200.trainer import TrainStep, synthetic_sample
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    torch.manual_seed(0)
    model = Model(names=NAMES, version=version).to(dev).train()
    model.train_backend = args.train_backend
    graph_x = torch.rand(per_gpu, 3, H, W, device=dev) if args.train_graphs else None
    ts = TrainStep(model, graph_sample=graph_x)
    K, Wm = args.steps, max(args.warmup, 3)
    host = [synthetic_sample(per_gpu, H, W, seed=7 * rank + i) for i in range(2)]
    host = [{k: v.pin_memory() for k, v in smp.items()} for smp in host]
    resident = [{k: v.to(dev) for k, v in smp.items()} for smp in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(Wm):
        ts.step(resident[i % 2])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        loss = ts.step(resident[i % 2])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join()
    # end to end: host sample -> H2D -> step -> loss to host
    Ke = min(K, 10)
    barrier()
    t0 = time.perf_counter()
    # the next sample's H2D copy runs on a copy stream while the current step computes (what a pinned-memory data loader does);
    # every step's copy and its loss read-back are inside the timed region
    copy_stream = torch.cuda.Stream(dev)

    def fetch(i):
        with torch.cuda.stream(copy_stream):
            smp = {k: v.to(dev, non_blocking=True) for k, v in host[i % 2].items()}
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        return smp, ev

    nxt = fetch(0)
    for i in range(Ke):
        smp, ev = nxt
        torch.cuda.current_stream(dev).wait_event(ev)
        for v in smp.values():
            v.record_stream(torch.cuda.current_stream(dev))
        if i + 1 < Ke:
            nxt = fetch(i + 1)
        lv = float(ts.step(smp))
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t[0]), float(t[1])
    if rank == 0:
        n_conv = sum(1 for m_ in model.modules() if isinstance(m_, torch.nn.Conv2d))
        # per sample: 2 forwards + their backward (dgrad + wgrad = 2x forward; the stem has no dgrad) = 6 x forward conv FLOPs
        flops_step = 6.0 * CONV_GFLOP[args.workload] * 1e9 * per_gpu
        step_s = ms * 1e-3 / K
        ach = flops_step / step_s / 1e12
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        line = {"metric": "training samples/sec (2 fwd + losses + bwd + grad all-reduce + Adam)", "value": K * per_gpu * world / (ms * 1e-3),
                "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16 operands, fp32 accumulation, fp32 master weights / BN / Adam", "data": "synthetic",
                "config": {"workload": f"YOLOPoint-{version.upper()} {W}x{H} training step, batch {per_gpu}/GPU (global {per_gpu * world})",
                           "conv_backend": args.train_backend, "cuda_graphs": bool(args.train_graphs), "parallelism": f"data parallel over {world} GPU(s), bucketed flat-buffer gradient all-reduce",
                           "l2": "activations of one step (GBs) exceed L2"},
                "clocks": sampler.summary(), "gpu_launches": K * n_conv * 2 * 3,
                "e2e": {"value": Ke * per_gpu * world / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "steps": Ke},
                "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel + wgrad_tc_kernel (tcgen05 conv forward / data gradient / weight gradient)",
                             "achieved": ach, "peak": peaks["tf"], "unit": "TFLOP/s", "frac": ach / peaks["tf"], "traffic": None,
                             "peak_source": f"{peaks['src']} bf16 sustained", "algorithmic_gflop_per_step": flops_step / 1e9,
                             "note": "whole-step time (convs, BN + SiLU, glue, losses, all-reduce, Adam); 6 x forward conv FLOPs per sample"},
                "detail": {"loss": lv}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="s640", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=1, help="independent camera streams (frame pipelines) in flight per GPU")
    ap.add_argument("--tile-policy", default="auto", choices=["auto", "latency", "wide"],
                    help="conv tile plan: latency = measured per-layer (tile_n, split_k) tables (fastest single pass), wide = widest N tile and no split-K "
                         "(fewest SM-seconds per pass: what several frames in flight want); auto = wide when --in-flight >= 4 at batch 1")
    ap.add_argument("--in-flight", type=int, default=0, help="frames of the one camera stream in flight per GPU (software pipelining; 1 = one frame at a time; 0 = auto: 8 at batch 1, 2 for batched workloads)")
    ap.add_argument("--also-streams", type=int, default=3, help="extra (untimed-for-value) run with this many camera streams, reported under detail")
    ap.add_argument("--train-backend", default="b200", choices=["b200", "cudnn_bf16", "torch"], help="l640train only: conv kernels used by the step")
    ap.add_argument("--train-batch", type=int, default=0, help="l640train only: samples per GPU (default 8)")
    ap.add_argument("--train-graphs", type=int, default=1, help="l640train only: 1 = forward/backward passes replayed from CUDA graphs, 0 = eager launches")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short runs of BASELINE configs 3-5 and of eager PyTorch reported under detail")
    args = ap.parse_args()
    version, H, W, per_gpu = WORKLOADS[args.workload]
    model_name = MODEL_NAME.get(args.workload, "YOLOPoint")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "l640train" and args.impl != "reference":
        return train_main(args, version, H, W, args.train_batch or per_gpu, world, rank, local_rank)
    K, Wm = args.steps, max(args.warmup, 3)
    config = {"workload": f"{model_name}-{version.upper()} {W}x{H} batch={per_gpu}/GPU full per-frame pipeline (net+decode+boxNMS+heatmap+kpNMS+desc+match)",
              "frames_per_gpu_per_step": per_gpu, "precision": args.precision, "parallelism": f"frames sharded over {world} GPU(s), no collective"}
    # execution parameters of the B200 arm; the reference arm prints the same `config` (it names the line it is compared with)
    if args.in_flight <= 0:
        args.in_flight = 8 if per_gpu == 1 else 2
    policy = args.tile_policy if args.tile_policy != "auto" else ("wide" if (args.in_flight >= 4 and per_gpu == 1) else "latency")
    NS = max(1, args.streams)
    frame_bytes = per_gpu * H * W * 3
    n_pool = max(8, int(160e6 // frame_bytes) + 1)
    # One step = one pipeline window = `frames_in_flight` consecutive frame batches of the camera stream (each its own batch-1 pass):
    # the contract's barrier + synchronize in front of the timed region empties the pipeline, so a step of a single frame would time
    # the fill / drain ramp of the window (K = 20: 27 frame slots for 20 frames), not the stream.
    FPS = args.in_flight
    config["frames_per_gpu_per_step"] = per_gpu * FPS
    config["step"] = f"{FPS} consecutive frame batch(es) of {per_gpu} frame(s): one pipeline window of the camera stream"
    config["streams_per_gpu"] = NS
    config["conv_tile_policy"] = policy + (" (persistent grids capped at a third of the SMs)" if (policy == "wide" and per_gpu == 1 and args.in_flight >= 4) else "")
    config["frames_in_flight"] = (f"{args.in_flight}: consecutive frames of the one camera stream are software-pipelined, each a batch-{per_gpu} pass; "
                                  "only the in-box filter + match of frame i+1 wait for frame i") if args.in_flight > 1 else 1
    config["l2"] = f"inputs larger than L2: pool of {n_pool} frame batches ({n_pool * frame_bytes / 1e6:.0f} MB) cycled, weights stay L2-resident"

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = min(K, 30)       # bounded sample: each CPU frame costs ~0.05-0.2 s
        warm = min(Wm, 10)
        fps, cores, med, sweep = cpu_reference_fps(version, H, W, steps, warm, model_name=model_name)
        line = {"impl": "reference", "metric": "frames/sec end-to-end (backbone+heads+NMS+match)", "value": fps, "unit": "frames/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                 "sample": f"{steps} timed frames ({warm} warm-up) of the same workload, batch 1, "
                                           f"oracle port of the reference CPU path; threads tuned over {sweep} frames/s"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    from yolopoint_b200 import FramePipeline
    from yolopoint_b200.synth import synthetic_frame
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    model, sd = build_weights(version, model_name)
    model.precision = args.precision
    model.tile_policy = policy
    model.wide_grid_div = 3 if (policy == "wide" and per_gpu == 1 and args.in_flight >= 4) else 1     # three layers of different frames share the GPU
    model = model.to(dev).eval()
    pipes = [FramePipeline(model, per_gpu, H, W, slot=i, frames_in_flight=args.in_flight) for i in range(NS)]
    cuda_streams = [torch.cuda.Stream(dev) for _ in range(NS)]
    for pp in pipes:
        pp.prepare()       # graph capture of every frame context (set-up; the W warm-up steps below then run the replay path)
    pipe = pipes[0]
    plan = pipe.plan

    # input pool larger than L2 (126 MB): every step reads a different frame batch from HBM
    # Weak scaling: every rank processes the SAME set of synthetic frames (identical work per GPU), each starting at a different
    # offset of the pool so that the ranks are not in lockstep.  (r01 seeded the frames by rank; the work per frame -- box
    # candidates 300 .. 7000 -- then differed between ranks, which measures the frames, not the scaling.  Frames of every rank
    # seed are covered by tests/test_gpu_capacity.py::test_bench_rank_seeds_do_not_raise.)
    base = [synthetic_frame(H, W, s) for s in range(4)]
    pool = torch.empty((n_pool, per_gpu, H, W, 3), dtype=torch.uint8, device=dev)
    for i in range(n_pool):
        for b in range(per_gpu):
            pool[i, b] = torch.from_numpy(np.roll(base[(i + b) % 4], shift=(3 * i) % W, axis=1)).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(i):
        """One step = every camera stream processes its next frame batch (independent graphs on independent CUDA streams)."""
        if NS == 1:
            for j in range(FPS):
                pipe.plan.frame_in.copy_(pool[(i * FPS + j + 7 * rank) % n_pool])
                pipe.step_device(True)
            return
        cur = torch.cuda.current_stream(dev)
        for s_i, (pp, cs) in enumerate(zip(pipes, cuda_streams)):
            cs.wait_stream(cur)
            with torch.cuda.stream(cs):
                for j in range(FPS):
                    pp.plan.frame_in.copy_(pool[((i * FPS + j) * NS + s_i + 7 * rank) % n_pool])
                    pp.step_device(True)
        for cs in cuda_streams:
            cur.wait_stream(cs)

    for i in range(Wm):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        step(Wm + i)
    for pp in pipes:
        pp.join()          # frames in flight on the pipelines' own streams end inside the timed region
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join()
    launches = K * FPS * NS * (pipe.n_launches())

    # ---- informational: several independent camera streams in flight (each batch 1), reported under detail only
    multi = None
    if NS == 1 and args.also_streams > 1:
        M = args.also_streams
        xp = [FramePipeline(model, per_gpu, H, W, slot=8 + i) for i in range(M)]
        xs = [torch.cuda.Stream(dev) for _ in range(M)]

        def mstep(i):
            cur = torch.cuda.current_stream(dev)
            for s_i, (pp, cs) in enumerate(zip(xp, xs)):
                cs.wait_stream(cur)
                with torch.cuda.stream(cs):
                    pp.plan.frame_in.copy_(pool[(i * M + s_i) % n_pool])
                    pp.step_device(True)
            for cs in xs:
                cur.wait_stream(cs)
        for i in range(4):
            mstep(i)
        torch.cuda.synchronize(dev)
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Km = min(K, 100)
        m0.record()
        for i in range(Km):
            mstep(i)
        m1.record()
        torch.cuda.synchronize(dev)
        multi = {"streams": M, "fps_per_gpu": Km * M * per_gpu / (m0.elapsed_time(m1) * 1e-3)}

    # ---- dominant kernel: conv launches only, timed live with events on the launching stream
    def net_only():
        plan.run_net()
    for _ in range(3):
        plan.graphed("net_only", net_only)
    torch.cuda.synchronize(dev)
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(20, min(K, 200))
    n0.record()
    for _ in range(reps):
        plan.graphed("net_only", net_only)
    n1.record()
    torch.cuda.synchronize(dev)
    net_ms = n0.elapsed_time(n1) / reps
    n_conv = len(model.engine().net.conv_ops())
    flops_step = CONV_GFLOP[args.workload] * 1e9 * per_gpu   # conv FLOPs of one frame batch
    # conv FLOPs of the K timed steps / duration of the timed region (CUDA events): with several frames in flight the conv launches of
    # different frames overlap, so their rate over the region is the honest figure (it also holds the non-conv kernels, all overlapped);
    # with one frame in flight it is the rate over the conv launches of one pass, timed alone (net_only_ms)
    achieved_tf = (flops_step * NS * K * FPS / (ms * 1e-3) / 1e12) if pipe.F > 1 else (flops_step / (net_ms * 1e-3) / 1e12)

    drained = policy == "wide" and args.precision == "fp32" and os.environ.get("YP_CONV_DRAIN", "1") != "0"
    conv_kernel_name = ("conv_tc_drain_kernel (tcgen05 implicit-GEMM conv, persistent grid, accumulators drained every 4 MMAs; csrc/conv_tc.cu)"
                        if drained else "conv_tc_kernel (tcgen05 implicit-GEMM conv; csrc/conv_tc.cu)")

    # ---- end to end through the public host API
    host_frames = [np.stack([np.roll(base[(i + b) % 4], (5 * i) % W, axis=1) for b in range(per_gpu)]) for i in range(8)]
    for pp in pipes:
        pp.reset_tracking()

    def host_submit(i):
        if NS == 1:
            pipe.submit_host(host_frames[i % 8])
            return
        for s_i, (pp, cs) in enumerate(zip(pipes, cuda_streams)):
            with torch.cuda.stream(cs):
                pp.submit_host(host_frames[(i * NS + s_i) % 8])

    def host_collect():
        return [pp.collect() for pp in pipes][0]

    for i in range(max(3, 2 * pipe.nctx)):     # every frame context captures its graphs before the timed region
        host_submit(i)
        host_collect()
    barrier()
    # One camera stream, frames in order; the host runs one frame ahead (stages frame i+1 into pinned memory and enqueues its
    # H2D + pipeline + D2H while frame i is on the GPU, then unpacks frame i).  Every step's H2D and D2H are inside the timed region.
    # The region is ~65 ms of host-driven work at the driver's flags, so one scheduling hiccup of one rank's host thread moves the
    # figure by tens of percent (round-2 record at N = 4); it is therefore run E2E_REPEATS times (barrier in front of each, every
    # repeat complete with its own H2D / D2H), each repeat is the MAX over ranks, and the MEDIAN repeat is reported (e2e.repeats_s
    # lists all of them).
    Ke = min(K * FPS, 400)             # frame batches of the end-to-end loop (the same K steps, bounded)
    ahead = max(1, args.in_flight)      # frames the host keeps submitted beyond the one it collects
    e2e_runs = []
    for _rep in range(E2E_REPEATS):
        barrier()
        t0 = time.perf_counter()
        for j in range(min(ahead, Ke)):
            host_submit(j)
        for i in range(Ke):
            if i + ahead < Ke:
                host_submit(i + ahead)
            res = host_collect()
        torch.cuda.synchronize(dev)
        e2e_runs.append(time.perf_counter() - t0)
    kp_n, box_n, match_n = res[0][0].shape[1], res[0][2].shape[0], res[0][3].shape[1]

    t = torch.tensor([ms] + e2e_runs, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_runs = float(t[0]), [float(v) for v in t[1:]]
    e2e_s = sorted(e2e_runs)[len(e2e_runs) // 2]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    fps = K * FPS * NS * per_gpu * world / (ms * 1e-3)
    line = {"metric": "frames/sec end-to-end (backbone+heads+NMS+match)", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (3xTF32 tensor-core MMA, fp32 accumulate)" if args.precision == "fp32" else "bf16",
            "data": "synthetic", "config": config, "clocks": sampler.summary(), "gpu_launches": launches,
            "e2e": {"value": Ke * NS * per_gpu * world / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": FPS * NS * pipe.h2d_bytes(),
                    "d2h_bytes_per_step": FPS * NS * pipe.d2h_bytes(), "steps": Ke / FPS,
                    "stat": f"median of {E2E_REPEATS} repeats of the same steps, each the max over ranks", "repeats_s": e2e_runs},
            "roofline": {"bound": "tensor", "kernel": conv_kernel_name, "achieved": achieved_tf, "peak": peaks["tf"],
                         "unit": "TFLOP/s", "frac": achieved_tf / peaks["tf"], "traffic": CONV_DRAM_BYTES_PER_LAUNCH.get(args.workload),
                         "traffic_note": "dram__bytes_read+write per conv launch, mean over the 64 conv launches of one pass of the headline plan (ncu, cold L2, serialised launches: "
                                         "profiles/r02_drain_pass_dram.md; --set full of 24 of them: profiles/r02_conv_tc_drain_ncu_full.md); algorithmic (tools/plan_bytes.py): 6.25 MB read "
                                         "(activations in (hi, lo) planes + weights, each once) + 4.04 MB written per launch; the measured reads are 1.03x the algorithmic reads, the writes "
                                         "stay in the write-back L2 beyond the end of the kernel and are read from there by the next layer",
                         "peak_source": f"{peaks['src']} bf16 sustained",
                         "launches_per_step": FPS * plan.n_net_launches(), "avg_launch_us": (ms / (K * FPS) if pipe.F > 1 else net_ms) * 1e3 / plan.n_net_launches(),
                         "algorithmic_gflop_per_step": FPS * flops_step / 1e9, "single_pass_ms": net_ms,
                         "note": ("achieved = SURVEY 8a conv FLOPs per frame x frames of the timed region / CUDA-event duration of the timed region (frames overlap: "
                                  "the conv launches of several frames run concurrently); single_pass_ms = the conv launches of ONE frame timed alone")
                                 if pipe.F > 1 else
                                 "achieved = SURVEY 8a conv FLOPs per frame x frames per step / live CUDA-event time of the conv launches of one step"},
            "detail": {"net_only_ms": net_ms, "keypoints": kp_n, "boxes": box_n, "matches": match_n, "concurrent_camera_streams": multi}}
    if world == 1 and not args.no_other_configs and args.workload == "s640":
        line["detail"]["torch_eager_gpu"] = torch_eager_net_ms(sd, version, model_name, per_gpu, H, W, dev)
        line["detail"]["other_configs"] = other_configs()
    if not args.no_cpu_baseline and world == 1:
        n = max(3, min(30, int(args.cpu_seconds / 0.3)))
        cfps, cores, med, sweep = cpu_reference_fps(version, H, W, n, 2, sd, model_name)
        line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"{n} frames of the same workload (batch 1), median {med * 1e3:.1f} ms/frame; threads tuned over {sweep} frames/s"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
