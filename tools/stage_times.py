#!/usr/bin/env python
"""Per-stage device times of resident steps under different cache states (diagnostic, not the benchmark)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from excord_lr_b200 import api, synth  # noqa: E402
from excord_lr_b200.batch import ExlrParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=1)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--cigar-kernel", type=int, default=0)
a = ap.parse_args()
c = synth.CONFIGS[a.config]
hb = synth.config(a.config, a.scale)
ex = api.Extractor(ExlrParams.make(**c["params"]), hb.ref_names, 0)
ex.set_option(api.EXLR_OPT_CIGAR_KERNEL, a.cigar_kernel)
b = ex.batch_for(hb)
b.upload()
flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
small = torch.zeros(8 << 20, dtype=torch.uint8, device="cuda")
for overlap in (0, 1):
    ex.set_option(api.EXLR_OPT_OVERLAP, overlap)
    for mode in ("flush256_read", "flush256_write", "touch8MB", "none"):
        acc = {}
        for i in range(a.steps + 3):
            if mode == "flush256_read":
                flush.sum()
            elif mode == "flush256_write":
                flush.fill_(1)
            elif mode == "touch8MB":
                small.sum()
            torch.cuda.synchronize()
            b.submit_resident()
            b.wait_resident()
            if i >= 3:
                for k, v in b.timing().as_dict().items():
                    if k.endswith("_ms"):
                        acc.setdefault(k, []).append(v * 1e3)
        print(f"overlap={overlap} {mode:15s}", {k[:-3]: round(float(np.median(v)), 1) for k, v in acc.items()}, flush=True)
