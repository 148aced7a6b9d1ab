"""One launch of every training glue / object-loss kernel at YOLOPoint-L 640x640 batch-8 sizes, for
`ncu --set full --clock-control none -k regex:"cat_|sppf_train|obj_" -o gpurun_out/r02_glue python tools/prof_glue.py`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from yolopoint_b200 import Model, losses as Lz, train as T  # noqa: E402

dev = torch.device("cuda")
CL = torch.channels_last
B = 8
rnd = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16).contiguous(memory_format=CL)

# cat(ups(xe), xb) at 80x80 and a C3 concat at 160x160, forward + backward
for parts, modes in (([rnd(B, 256, 40, 40), rnd(B, 256, 80, 80)], ["up2", "copy"]), ([rnd(B, 64, 160, 160), rnd(B, 64, 160, 160)], ["copy", "copy"])):
    ps = [p.requires_grad_(True) for p in parts]
    out = T.cat_tc(ps, modes)
    out.backward(torch.ones_like(out))
# SPPF cascade
x = rnd(B, 512, 20, 20).requires_grad_(True)
o = T.sppf_cat_tc(x)
o.backward(torch.ones_like(o))
# object loss, 64 targets, in-kernel target assignment
torch.manual_seed(0)
m = Model(names=[str(i) for i in range(80)], version="n").to(dev)
cfg = dict(box=0.05, cls=0.5, cls_pw=1.0, obj=1.0, obj_pw=1.0, iou_t=0.2, anchor_t=4.0, label_smoothing=0.0, fl_gamma=0.0)
crit = Lz.ComputeObjectLoss(m, cfg, dev)
gen = torch.Generator().manual_seed(1)
nt = 64
tg = torch.cat((torch.randint(0, B, (nt, 1), generator=gen).float(), torch.randint(0, 80, (nt, 1), generator=gen).float(),
                0.1 + 0.8 * torch.rand(nt, 2, generator=gen), 0.05 + 0.35 * torch.rand(nt, 2, generator=gen)), 1).to(dev)
p = [torch.randn(B, 3, s, s, 85, device=dev).requires_grad_(True) for s in (80, 40, 20)]
loss, _ = crit(p, tg)
loss.sum().backward()
torch.cuda.synchronize()
print("done", float(loss))
