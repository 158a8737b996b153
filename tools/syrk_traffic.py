"""DRAM traffic / throughput of the d = 14336 SYRK launch against tokens per launch and raster band.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:gemm_tc2 --csv python tools/syrk_traffic.py --once      # one launch per config
    python tools/syrk_traffic.py                                          # sustained TF/s per config
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from ptdeco_b200 import _native as nat
from ptdeco_b200 import linalg
from tools.syrk_bench import sustained

CONFIGS = [(8192, 8), (8192, 12), (16384, 4), (16384, 6), (16384, 8), (12288, 6)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--d", type=int, default=14336)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    d = args.d
    L = nat.lib()
    g = torch.Generator(device=dev).manual_seed(d)
    y = torch.randn(16384, d, generator=g, device=dev).to(torch.bfloat16)
    acc = linalg.CovarianceAccumulator(d, dev)
    for (n, band) in CONFIGS:
        L.ptdeco_debug_set(9, band)
        yy = y[:n]
        if args.once:
            acc._syrk(yy, None, 1.0 / n)
            torch.cuda.synchronize()
            print(json.dumps({"N": n, "band": band}), flush=True)
        else:
            rec = {"N": n, "band": band}
            rec.update(sustained(lambda: acc._syrk(yy, None, 1.0 / n), n * d * (d + 1.0), 1.2))
            print(json.dumps(rec), flush=True)
    L.ptdeco_debug_set(9, 0)


if __name__ == "__main__":
    main()
