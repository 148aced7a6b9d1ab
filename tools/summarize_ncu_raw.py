"""Markdown table from an `ncu -i X.ncu-rep --page raw --csv` export: python tools/summarize_ncu_raw.py raw.csv [max_rows_per_kernel_shape]
Columns: duration, DRAM bytes, achieved DRAM GB/s and % of the measured HBM peak, tensor-pipe activity, SM throughput, regs, smem.
Tensor-pipe activity: `sm__pipe_tensor_cycles_active*.pct` does not see tcgen05 (UTCHMMA) work; the counter that moves is
`sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg` (cycles per SM with an MMA in the tensor sub-pipe), reported here as a
fraction of `sm__cycles_elapsed.avg` over all SMs and scaled to the SMs the grid can occupy."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
h, data = rows[0], rows[2:]
ix = {k: i for i, k in enumerate(h)}
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6550.0}


def f(r, k, d=0.0):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return d


TENS = "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"
HMMA = "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"
print(f"| kernel | grid | us | DRAM read MB | DRAM write MB | DRAM GB/s | % of HBM peak ({peaks['hbm_gbs']:.0f} GB/s) | tensor pipe active % (pct counter) | "
      f"tensor sub-pipe (hmma) cycles active, % of elapsed: all SMs / busy SMs | SM throughput % | regs | dyn smem KB |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
seen = {}
for r in data:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("yp::<", "").strip()
    key = (name, r[ix["Grid Size"]], r[ix["launch__shared_mem_per_block_dynamic"]])
    seen[key] = seen.get(key, 0) + 1
    if seen[key] > 1:
        continue
    us = f(r, "gpu__time_duration.sum")
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    units = rows[1][ix["dram__bytes_read.sum"]]
    scale = {"Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}.get(units, 1.0)
    rd, wr = rd * scale, wr * scale
    gbs = (rd + wr) * 1e6 / (us * 1e-6) / 1e9 if us else 0.0
    ctas = 1
    for v in r[ix["Grid Size"]].strip("() ").split(","):
        ctas *= int(v)
    el = f(r, "sm__cycles_elapsed.avg", 1.0) or 1.0
    hm = 100.0 * f(r, HMMA) / el if HMMA in ix else 0.0
    print(f"| `{name}` | {r[ix['Grid Size']]} | {us:.1f} | {rd:.2f} | {wr:.2f} | {gbs:.0f} | {100 * gbs / peaks['hbm_gbs']:.1f} | {f(r, TENS):.1f} | "
          f"{hm:.1f} / {min(100.0, hm * 148.0 / min(ctas, 148)):.1f} | "
          f"{f(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {int(f(r, 'launch__registers_per_thread'))} | {f(r, 'launch__shared_mem_per_block_dynamic'):.0f} |")
