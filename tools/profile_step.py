#!/usr/bin/env python
"""A few device-resident steps of one BASELINE config, for ncu to attach to (not a benchmark).
   ncu --set full --import-source on --clock-control none -k regex:'k3b|k1a|k1_flat' --launch-skip 6 -c 3 -o gpurun_out/x python tools/profile_step.py"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from excord_lr_b200 import api, synth  # noqa: E402
from excord_lr_b200.batch import ExlrParams  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=1)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--cigar-kernel", type=int, default=0)
ap.add_argument("--overlap", type=int, default=0)
ap.add_argument("--graph", type=int, default=1)
a = ap.parse_args()
c = synth.CONFIGS[a.config]
hb = synth.config(a.config, a.scale)
ex = api.Extractor(ExlrParams.make(**c["params"]), hb.ref_names, 0)
ex.set_option(api.EXLR_OPT_CIGAR_KERNEL, a.cigar_kernel)
ex.set_option(api.EXLR_OPT_OVERLAP, a.overlap)
ex.set_option(api.EXLR_OPT_STAGE_TIMING, 0)
ex.set_option(api.EXLR_OPT_GRAPH, a.graph)
b = ex.batch_for(hb)
b.upload()
flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(a.steps):
    flush.sum()
    torch.cuda.synchronize()
    b.submit_resident()
    r = b.wait_resident()
print("status", r.status, "lines", r.n_events)
