"""One training step of the reference (src/train.py:196-250) on the B200 path, data parallel over the GPUs of one box.

  step = forward(image) + forward(warped image)                    (tcgen05 convs, train.py)
       + object loss + 2 x detector loss + InfoNCE descriptor loss  (losses.py; train.py:8, 212-241 of the reference)
       + backward                                                   (tcgen05 dgrad / wgrad)
       + gradient all-reduce (mean over ranks)                      (NCCL over NVLink; gloo in the CPU tests)
       + Adam step (lr 1e-3 over all parameters, src/train.py:88) and the linear LambdaLR schedule (:91-93)

Data parallelism follows the reference's DDP setup (src/train.py:46: ``broadcast_buffers=False``, no SyncBN): every rank owns
its slice of the batch and its own BN statistics; gradients are averaged.  The gradients live in ONE flat fp32 buffer (each
``p.grad`` is a view of it), cut into buckets in reverse parameter order; a bucket is all-reduced asynchronously as soon as
autograd has produced its last gradient, so the collective overlaps the rest of the backward pass.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import os

import numpy as np
import torch
import torch.distributed as dist

from . import losses as Lz

LOSS_CFG = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, iou_t=0.2, anchor_t=4.0, label_smoothing=0.0, fl_gamma=0.0)  # configs/coco.yaml:128-140
SPARSE_CFG = dict(num_samples_per_image=3000, num_masked_non_matches_per_match=200)                                            # configs/coco.yaml:122-124
LAMBDA_DESC, LAMBDA_OBJ = 0.1, 10.0                                                                                             # configs/coco.yaml:113-115


class FlatGradReducer:
    """Flat gradient buffer + bucketed asynchronous all-reduce (mean)."""

    def __init__(self, params: List[torch.nn.Parameter], bucket_bytes: int = 32 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        # reverse parameter order ~ the order in which backward produces gradients: bucket 0 = the last parameters
        off = total
        self.bucket_of: Dict[int, int] = {}
        self.buckets: List[List[int]] = [[total, total, 0]]      # [lo, hi, n_params]
        for i in reversed(range(len(self.params))):
            p = self.params[i]
            off -= p.numel()
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            b = self.buckets[-1]
            if (b[1] - b[0]) * 4 >= bucket_bytes:
                self.buckets.append([off + p.numel(), off + p.numel(), 0])
                b = self.buckets[-1]
            b[0] = off
            b[2] += 1
            self.bucket_of[i] = len(self.buckets) - 1
        self.pending = [0] * len(self.buckets)
        self.handles = []
        if self.world > 1:
            for i, p in enumerate(self.params):
                p.register_post_accumulate_grad_hook(lambda _p, i=i: self._ready(i))

    def zero(self):
        self.flat.zero_()
        for i, p in enumerate(self.params):      # autograd accumulates in place into the existing views
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * self._offset(i):
                raise RuntimeError("a parameter's .grad no longer aliases the flat gradient buffer (do not call zero_grad(set_to_none=True))")
        self.pending = [b[2] for b in self.buckets]
        self.handles = []

    def _offset(self, i):
        if not hasattr(self, "_offs"):
            offs, o = [], 0
            for p in self.params:
                offs.append(o)
                o += p.numel()
            self._offs = offs
        return self._offs[i]

    def _ready(self, i):
        b = self.bucket_of[i]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            lo, hi, _ = self.buckets[b]
            self.handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Wait for the bucket all-reduces (buckets whose parameters received no gradient are reduced here) and average."""
        if self.world == 1:
            return
        for b, n in enumerate(self.pending):
            if n > 0:
                lo, hi, _ = self.buckets[b]
                self.handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for h in self.handles:
            h.wait()
        self.flat.mul_(1.0 / self.world)


def synthetic_sample(B: int, H: int, W: int, seed: int, device="cpu", boxes_per_image: int = 8, nc: int = 80) -> Dict[str, torch.Tensor]:
    """The synthetic training sample of SURVEY.md section 8d (config 5): images U(0,1), keypoint labels Bernoulli(0.005), valid masks
    of ones, 8 boxes per image, identity homographies."""
    rs = np.random.RandomState(seed)
    f = lambda a: torch.from_numpy(a.astype(np.float32)).to(device)
    boxes = np.concatenate([np.stack([np.full(boxes_per_image, b), rs.randint(0, nc, boxes_per_image), rs.uniform(0.1, 0.9, boxes_per_image),
                                      rs.uniform(0.1, 0.9, boxes_per_image), rs.uniform(0.05, 0.4, boxes_per_image), rs.uniform(0.05, 0.4, boxes_per_image)], 1)
                            for b in range(B)])
    return dict(image=f(rs.rand(B, 3, H, W)), warped_image=f(rs.rand(B, 3, H, W)), labels_2D=f(rs.rand(B, 1, H, W) < 0.005),
                warped_labels=f(rs.rand(B, 1, H, W) < 0.005), valid_mask=torch.ones(B, 1, H, W, device=device),
                warped_valid_mask=torch.ones(B, 1, H, W, device=device), box_labels=f(boxes), inv_homographies=torch.eye(3, device=device).repeat(B, 1, 1))


class _FlatOutputs(torch.nn.Module):
    """Model.forward with its train-mode outputs flattened to a tuple of tensors (what torch.cuda.make_graphed_callables wants)."""

    def __init__(self, model):
        super().__init__()
        self.m = model

    def forward(self, x):
        out = self.m(x)
        return (out["semi"], out["desc"], *out["objects"])


class _LossHead(torch.nn.Module):
    """The three losses of a step as one callable over tensors only (network outputs of both passes, labels, the precomputed target
    plan and descriptor pairs), so that torch.cuda.make_graphed_callables can capture its forward and backward: ~700 small kernels
    (target gather, CIoU, BCE terms, sampling, similarity GEMM, log-softmax and their gradients) replay from two graph launches
    instead of being issued one by one from Python."""

    def __init__(self, ts: "TrainStep", cells):
        super().__init__()
        self.ts, self.cells = [ts], list(cells)      # (list: keep the TrainStep out of nn.Module's attribute registry)

    def forward(self, semi, semi_w, desc, desc_w, p0, p1, p2, labels_2D, valid_mask, warped_labels, warped_valid_mask, inv_h, pa, pb, rnd, *plan):
        ts = self.ts[0]
        if len(plan) == 1:          # the label list itself: the target assignment runs inside the object-loss kernels
            loss_obj, _ = ts.obj_loss([p0, p1, p2], plan[0])
        else:
            levels = [dict(valid=plan[5 * i], cell=plan[5 * i + 1], tbox=plan[5 * i + 2], anchor=plan[5 * i + 3], cls=plan[5 * i + 4], cells=self.cells[i])
                      for i in range(len(self.cells))]
            loss_obj, _ = ts.obj_loss([p0, p1, p2], None, Lz.TargetPlan(levels))
        loss_det = ts.det_loss.from_2d(semi, labels_2D, valid_mask)
        loss_det_w = ts.det_loss.from_2d(semi_w, warped_labels, warped_valid_mask)
        loss_desc = ts.desc_loss(desc, desc_w, warped_valid_mask, inv_h, pairs=(pa, pb, rnd), **ts.sparse_cfg)
        loss = loss_det + loss_det_w + LAMBDA_DESC * loss_desc + LAMBDA_OBJ * loss_obj
        return loss.reshape(1), torch.stack((loss_det.detach().reshape(()), loss_det_w.detach().reshape(()), loss_desc.detach().reshape(()),
                                             loss_obj.detach().reshape(())))


class TrainStep:
    def __init__(self, model, epochs: int = 100, lr: float = 1e-3, lrf: float = 0.01, sparse_cfg: Optional[dict] = None, group=None,
                 bucket_bytes: int = 32 << 20, graph_sample: Optional[torch.Tensor] = None, desc_loss: str = "infonce",
                 gradclip: Optional[float] = None):
        """``graph_sample``: an image batch [B,3,H,W] on the model's device.  When given, the two forward passes of a step and
        their backward passes are captured into CUDA graphs (torch.cuda.make_graphed_callables: one forward + one backward graph
        per pass, shared parameters), so that the ~10^4 kernel launches of a step replay from four graph launches instead of
        being issued one by one from Python; the batch shape is then fixed.  BatchNorm statistics, the losses, the gradient
        all-reduce and Adam are unchanged.  ``desc_loss``: "infonce" = the descriptor loss of the reference's training script
        (src/train.py:8 imports ``infonce`` under the name ``descriptor_loss_sparse``), "hinge" = the older
        ``descriptor_loss_sparse`` of src/utils/loss_functions.py:361-481 (same sampled pairs, same similarities)."""
        if desc_loss not in ("infonce", "hinge"):
            raise ValueError(f"desc_loss={desc_loss!r}: expected 'infonce' or 'hinge'")
        self.desc_loss = Lz.infonce if desc_loss == "infonce" else Lz.descriptor_loss_sparse
        self.gradclip = gradclip                # src/train.py:249-250: clip_grad_norm_ on the (averaged) gradients before the optimizer step
        self.model = model
        self.device = next(model.parameters()).device
        # (the sampled-similarity GEMM of the descriptor loss runs on TF32 tensor cores: losses._pair_similarities switches
        # torch.backends.cuda.matmul.allow_tf32 on around that one matmul and restores it; nothing global is changed here)
        self.obj_loss = Lz.ComputeObjectLoss(model, LOSS_CFG, self.device)
        self.det_loss = Lz.ComputeDetectorLoss(self.device)
        self.sparse_cfg = dict(SPARSE_CFG if sparse_cfg is None else sparse_cfg)
        self.reducer = FlatGradReducer(list(model.parameters()), bucket_bytes, group)
        # (the foreach form, not fused=True: the fused kernel updates the parameters without advancing their version counters, which
        # the operand-pack cache of train.py keys on -- eager steps would then run on stale bf16 weight copies)
        self.opt = torch.optim.Adam(model.parameters(), lr=lr)
        self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lr_lambda=lambda e: (1 - e / epochs) * (1.0 - lrf) + lrf)
        self.graphed = None
        self.loss_heads = {}                    # shape key -> graphed _LossHead (or None when the capture failed: eager losses)
        self.graph_losses = graph_sample is not None and os.environ.get("YP_TRAIN_GRAPH_LOSSES", "1") != "0"
        self.packs, self.repack_graph = [], None
        self._bn_counts = []
        if graph_sample is not None:
            model.train()
            if getattr(model, "train_backend", None) == "b200":
                # operand copies of the weights live in persistent buffers refreshed by one small graph per step, so that the
                # graphs of the passes contain no packing kernels
                from . import train as _train
                _train.enable(model)
                model._tc_train = "b200"
                # num_batches_tracked of all BatchNorms: one foreach add per step instead of one tiny kernel per BatchNorm per pass
                for mod in model.modules():
                    if _train.fused_bn_block(mod) and mod.bn.num_batches_tracked is not None:
                        mod.bn._yp_defer_count = True
                        self._bn_counts.append(mod.bn.num_batches_tracked)
                self.packs = _train.attach_weight_packs(model)
                for pk in self.packs:
                    pk.refresh()
                torch.cuda.synchronize(self.device)
                self.repack_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.repack_graph):
                    for pk in self.packs:
                        pk.refresh()
            bn_state = {k: v.clone() for k, v in model.state_dict().items() if "running_" in k or "num_batches" in k}
            passes = (_FlatOutputs(model), _FlatOutputs(model))
            x = graph_sample.detach().clone()
            self.graphed = torch.cuda.make_graphed_callables(passes, ((x,), (x.clone(),)), num_warmup_iters=3)
            model.load_state_dict(bn_state, strict=False)      # the warm-up / capture passes updated the running statistics
            self.reducer.zero()

    def _forward(self, x, which):
        if self.graphed is None:
            out = self.model(x)
            return out["semi"], out["desc"], out["objects"]
        flat = self.graphed[which](x)
        return flat[0], flat[1], list(flat[2:])

    def losses(self, sample):
        m, dev = self.model, self.device
        # Everything that depends on the labels only and synchronises with the host (target assignment, sample-pool sizes) runs
        # first, so that the forward passes and the loss arithmetic can be enqueued back to back without the host waiting on them.
        B, _, H, W = sample["image"].shape
        det = getattr(m, "module", m).model.Detect
        shapes = [torch.empty((B, det.na, H // int(s), W // int(s), det.no), device="meta") for s in (8, 16, 32)]
        # (CUDA: the object-loss kernels assign the targets themselves from the label list; the PyTorch statement needs the plan)
        built = None if (self.obj_loss.fused and dev.type == "cuda") else self.obj_loss.build_targets(shapes, sample["box_labels"])
        pairs = Lz.descriptor_pairs(sample["warped_valid_mask"], sample["inv_homographies"], B, H // 8, W // 8, device=dev, **self.sparse_cfg)
        semi, desc, obj = self._forward(sample["image"], 0)
        semi_w, desc_w, _ = self._forward(sample["warped_image"], 1)
        if getattr(self, "graph_losses", False) and len(obj) == 3:
            res = self._graphed_losses(sample, semi, semi_w, desc, desc_w, obj, built, pairs)
            if res is not None:
                return res
        loss_obj, items = self.obj_loss(obj, sample["box_labels"], built)
        # label layout + cell masks + loss + gradient of the detector loss: one fused kernel per pass on CUDA (csrc/loss.cu)
        loss_det = self.det_loss.from_2d(semi, sample["labels_2D"], sample["valid_mask"])
        loss_det_w = self.det_loss.from_2d(semi_w, sample["warped_labels"], sample["warped_valid_mask"])
        loss_desc = self.desc_loss(desc, desc_w, sample["warped_valid_mask"], sample["inv_homographies"], pairs=pairs, **self.sparse_cfg)
        loss = loss_det + loss_det_w + LAMBDA_DESC * loss_desc + LAMBDA_OBJ * loss_obj
        return loss, dict(det=loss_det.detach(), det_warp=loss_det_w.detach(), desc=loss_desc.detach(), obj=loss_obj.detach())

    def _graphed_losses(self, sample, semi, semi_w, desc, desc_w, obj, built, pairs):
        """The loss head through a CUDA graph captured per shape signature (number of targets, size of the descriptor sample pool,
        batch geometry): a data loader delivers a handful of distinct signatures when its label lists are padded to a bucket size;
        a signature whose capture fails runs the eager losses (returns None)."""
        dev = self.device
        f32 = lambda t: t.to(dev).float().contiguous()
        if built is None:
            plan = [f32(sample["box_labels"])]
        else:
            plan = [lv[k].contiguous() for lv in built.levels for k in ("valid", "cell", "tbox", "anchor", "cls")]
        args = [semi, semi_w, desc, desc_w, obj[0], obj[1], obj[2], f32(sample["labels_2D"]), f32(sample["valid_mask"]), f32(sample["warped_labels"]),
                f32(sample["warped_valid_mask"]), f32(sample["inv_homographies"]), pairs[0].contiguous(), pairs[1].contiguous(), pairs[2].contiguous()] + plan
        key = tuple((tuple(t.shape), t.dtype) for t in args)
        if key not in self.loss_heads:
            head = _LossHead(self, [o.numel() // o.shape[-1] for o in obj])
            try:
                samples = tuple(t.detach().clone().requires_grad_(t.requires_grad) for t in args)
                self.loss_heads[key] = torch.cuda.make_graphed_callables(head, samples, num_warmup_iters=2)
            except Exception as e:      # pragma: no cover  (e.g. an op of a custom loss configuration that cannot be captured)
                import warnings
                warnings.warn(f"TrainStep: loss head not captured ({type(e).__name__}: {e}); running the losses eagerly")
                self.loss_heads[key] = None
        head = self.loss_heads[key]
        if head is None:
            return None
        loss, parts = head(*args)
        return loss[0], dict(det=parts[0], det_warp=parts[1], desc=parts[2], obj=parts[3])

    def epoch_end(self):
        """Advance the linear learning-rate schedule (src/train.py:91-93, stepped once per epoch at :289): the caller's epoch loop
        calls this; ``step`` itself never touches the schedule."""
        self.sched.step()

    def step(self, sample) -> torch.Tensor:
        """sample: dict as produced by the reference's data loader (src/train.py:196-205) with tensors on the model's device."""
        self.model.train()
        self.reducer.zero()
        if self.repack_graph is not None:
            self.repack_graph.replay()          # bf16 operand copies of the weights the optimizer just updated
        loss, _ = self.losses(sample)
        if self._bn_counts:
            torch._foreach_add_(self._bn_counts, 2)     # two forward passes per step (frame and warped frame)
        loss.backward()
        self.reducer.finish()
        if self.gradclip:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.gradclip)
        self.opt.step()
        return loss.detach()
