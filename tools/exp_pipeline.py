#!/usr/bin/env python
"""Experiment: device-resident steps of SEVERAL batches in flight (each on its own streams) against one batch stepped alone.
   python tools/exp_pipeline.py [config] [n_batches ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from excord_lr_b200 import api, synth  # noqa: E402
from excord_lr_b200.batch import ExlrParams  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
counts = [int(x) for x in sys.argv[2:]] or [1, 2, 3, 4]
c = synth.CONFIGS[cfg]
hb = synth.config(cfg, 1.0)
ex = api.Extractor(ExlrParams.make(**c["params"]), hb.ref_names, 0)
ex.set_option(api.EXLR_OPT_STAGE_TIMING, 0)
for kv in os.environ.get("EXLR_EXP_OPTS", "").split(","):          # e.g. EXLR_EXP_OPTS=9:6  (option 9 = k1a CTAs per SM)
    if kv:
        ex.set_option(int(kv.split(":")[0]), int(kv.split(":")[1]))
flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
for nb in counts:
    bs = [ex.batch_for(hb) for _ in range(nb)]
    for b in bs:
        b.upload()
    def run(k):
        busy = [False] * nb
        for s in range(k):
            i = s % nb
            if busy[i]:
                r = bs[i].wait_resident(); assert r.status == 0
            bs[i].submit_resident(); busy[i] = True
        for i in range(nb):
            if busy[i]:
                r = bs[i].wait_resident(); assert r.status == 0
        return r.n_events
    run(12)
    best = None
    for _ in range(5):
        flush.sum(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        n = run(60)
        e1.record(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f"config {cfg}: {nb} batch(es) in flight: {best / 60 * 1e6:.1f} us per step (wall, 60 steps, best of 5; events {e0.elapsed_time(e1) / 60 * 1e3:.1f} us), {hb.n_reads / (best / 60):.3e} alignments/s, lines {n}")
    for b in bs:
        b.free()
