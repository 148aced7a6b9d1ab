"""Eager network passes of YOLOPoint-S 640x640 batch 1 for an ncu capture of every conv launch under a tile policy:
   ncu --set full --clock-control none -k regex:conv_tc_kernel -s 130 -c 65 -o gpurun_out/x python tools/prof_net_pass.py [latency|wide] [grid divisor of the wide plan]
   per-launch DRAM traffic of a whole pass (a few replays per kernel instead of the ~40 of --set full):
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none
       -k regex:conv_tc --csv --log-file gpurun_out/x.csv python tools/prof_net_pass.py wide 3"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from yolopoint_b200 import Model  # noqa: E402
from yolopoint_b200.engine import Engine  # noqa: E402
from yolopoint_b200.synth import perturb_state_dict  # noqa: E402

policy = sys.argv[1] if len(sys.argv) > 1 else "wide"
grid_div = int(sys.argv[2]) if len(sys.argv) > 2 else None      # bench.py's headline configuration: wide 3
torch.manual_seed(0)
m = Model(names=[str(i) for i in range(80)], version="s")
sd = perturb_state_dict(m.state_dict(), 0, "s")
eng = Engine(sd, "s", 80, torch.device("cuda:0"), tile_policy=policy, use_graphs=False, multi_stream=False, wide_grid_div=grid_div)
p = eng.plan(1, 640, 640)
p.frame_in.copy_(torch.randint(0, 256, p.frame_in.shape, dtype=torch.uint8, device="cuda"))
for _ in range(4):
    p.run_input(True)
    p.run_net()
torch.cuda.synchronize()
print("conv launches per pass:", len(eng.net.conv_ops()))
