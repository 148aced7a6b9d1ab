"""Layer chains (yp_conv_chain_*) against the per-layer launch list on a B200: same outputs, time per network pass.

usage: python tools/chain_check.py [version] [B] [H] [W]     (default: s 1 640 640, the headline configuration)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolopoint_b200 import Model  # noqa: E402
from yolopoint_b200.engine import Engine  # noqa: E402
from yolopoint_b200.synth import perturb_state_dict  # noqa: E402


def main():
    ver = sys.argv[1] if len(sys.argv) > 1 else "s"
    B, H, W = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (1, 640, 640)
    torch.manual_seed(0)
    m = Model(names=[str(i) for i in range(80)], version=ver)
    sd = perturb_state_dict(m.state_dict(), 0, ver)
    dev = torch.device("cuda:0")
    x = torch.rand(B, 3, H, W, device=dev)
    res = {}
    for chain in (False, True):
        eng = Engine(sd, ver, 80, dev, chain=chain)
        p = eng.plan(B, H, W)
        out = eng.forward(x)
        torch.cuda.synchronize()
        out2 = eng.forward(x)   # second launch: the completion counters must have been reset
        torch.cuda.synchronize()
        assert torch.equal(out["semi"], out2["semi"]) and torch.equal(out["objects"][0], out2["objects"][0]), "replay differs"
        for _ in range(3):
            p.graphed("net_only", p.run_net)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            p.graphed("net_only", p.run_net)
        e1.record()
        torch.cuda.synchronize()
        res[chain] = dict(out=out, ms=e0.elapsed_time(e1) / 200, launches=p.n_net_launches(),
                          segments=[(len(s["ops"]), s["kernels"], s["items"], s["smem"]) for s in (p.chain or [])])
        if chain and p.chain:
            # every segment alone (its inputs are still in the buffers) next to the same ops as single-stream per-layer launches
            import ctypes as C
            from yolopoint_b200 import _lib
            L = _lib.lib()
            seg_ms = []
            for sg in p.chain:
                def run_seg(sg=sg):
                    _lib.check(L.yp_conv_chain_launch(sg["handle"], C.c_void_p(torch.cuda.current_stream().cuda_stream)))
                def run_ops(sg=sg):
                    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                    for j in sg["ops"]:
                        p.launches[j][1](st)
                t = []
                for fn in (run_seg, run_ops):
                    g = torch.cuda.CUDAGraph()
                    fn(); torch.cuda.synchronize()
                    with torch.cuda.graph(g):
                        fn()
                    for _ in range(3):
                        g.replay()
                    e0.record()
                    for _ in range(100):
                        g.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    t.append(round(e0.elapsed_time(e1) * 10, 1))   # us per replay
                seg_ms.append(t)
            res[chain]["segment_us(chain,launches)"] = seg_ms
    a, b = res[False]["out"], res[True]["out"]
    err = {k: float((a[k] - b[k]).abs().max()) / max(1.0, float(a[k].abs().max())) for k in ("semi", "desc")}
    err["pred"] = float(((a["objects"][0] - b["objects"][0]).abs() / (1.0 + a["objects"][0].abs())).max())
    print(json.dumps({"config": [ver, B, H, W], "ms_per_layer_launches": res[False]["ms"], "ms_chain": res[True]["ms"],
                      "launches": [res[False]["launches"], res[True]["launches"]], "segments(ops,kernels,items,smem)": res[True]["segments"], "segment_us(chain,launches)": res[True].get("segment_us(chain,launches)"),
                      "max_rel_diff_chain_vs_launches": err}))


if __name__ == "__main__":
    main()
