#!/usr/bin/env python
"""Wall-clock of the full BAM -> event-file path through the CLI (excord_lr_b200/host/excord-lr-b200 --stats), on BAMs written
from the synthetic configs.  Reports honestly where the time goes: BGZF inflate + record walk + packing (host), GPU, formatting.

  python tools/cli_timing.py [--threads 16]
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from excord_lr_b200 import bamio, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
args = ap.parse_args()
exe = os.path.join(ROOT, "excord_lr_b200", "host", "excord-lr-b200")
subprocess.check_call(["make", "-C", os.path.join(ROOT, "excord_lr_b200", "host")], stdout=subprocess.DEVNULL)

cases = [("c2 HiFi 1M molecules, no SEQ/QUAL in the BAM", 1, 1.0, 0), ("c2 HiFi 30k molecules with 15 kb random SEQ/QUAL per record", 1, 0.03, 15000)]
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    for name, cfg, scale, seq_len in cases:
        hb = synth.config(cfg, scale)
        bam, out = os.path.join(d, "x.bam"), os.path.join(d, "x.txt")
        t0 = time.time()
        bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=seq_len, random_seq=seq_len > 0)
        tw = time.time() - t0
        size = os.path.getsize(bam)
        best = None
        for _ in range(3):
            t0 = time.time()
            r = subprocess.run([exe, "-b", bam, "-o", out, "-p", "0.8", "-t", str(args.threads), "--stats"], capture_output=True, text=True)
            dt = time.time() - t0
            assert r.returncode == 0, r.stderr
            best = dt if best is None or dt < best else best
            stats = r.stderr.strip().splitlines()[-1]
        print(f"{name}: {hb.n_reads} records, BAM {size / 1e6:.0f} MB (written in {tw:.0f} s), {args.threads} inflate threads")
        print(f"  best of 3 wall {best:.3f} s -> {hb.n_reads / best:.3e} alignments/s, {size / best / 1e6:.0f} MB/s of BAM")
        print(f"  {stats}")
