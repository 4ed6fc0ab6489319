#!/usr/bin/env python
"""Writes a synthetic BAM of a BASELINE.json config (for CLI timing / ncu runs on the GPU box).
  python tools/make_bam.py out.bam [--config 1] [--scale 0.03] [--seq-len 15000] [--level 1]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from excord_lr_b200 import bamio, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("out")
ap.add_argument("--config", type=int, default=1)
ap.add_argument("--scale", type=float, default=0.03)
ap.add_argument("--seq-len", type=int, default=15000)
ap.add_argument("--level", type=int, default=1)
a = ap.parse_args()
hb = synth.config(a.config, a.scale)
bamio.write_bam(hb, a.out, ref_lens=synth.ref_lens(), seq_len=a.seq_len, random_seq=a.seq_len > 0, level=a.level)
print(f"{a.out}: {hb.n_reads} records, {os.path.getsize(a.out) / 1e6:.1f} MB")
