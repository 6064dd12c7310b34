"""Compact summaries of ncu outputs for profiles/ (run on the CPU box).

    python tools/summarize_ncu.py launches gpurun_out/x_launches.csv   # per-kernel share of ONE step (last im2col -> end)
    python tools/summarize_ncu.py full gpurun_out/x_prof.ncu-rep       # key metrics per captured launch
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_xu.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def short(name):
    name = name.replace("void ", "").replace("vb::", "")
    return name.split("(")[0][:70]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        seq.append((short(r[ki]), us))
    # one step = from the last im2col launch to the end of the list
    starts = [i for i, (k, _) in enumerate(seq) if "im2col" in k]
    step = seq[starts[-1]:] if starts else seq
    agg = OrderedDict()
    for k, us in step:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list, one finetune step ({len(step)} launches of {len(seq)} captured); per-launch times are cold-cache + serialised: compare SHARES")
    print(f"# total {tot/1e3:.2f} ms")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us/1e3:9.3f} ms {100*us/tot:5.1f}%  x{n:<4d} {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"## {short(r[ki])}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:95s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
