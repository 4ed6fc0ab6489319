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
        print(f"{name}: {hb.n_reads} records, BAM {size / 1e6:.0f} MB (written in {tw:.0f} s)")
        ref_out = None
        for mode, extra in (("GPU BAM decoder (default)", []), (f"host reader, {args.threads} zlib threads (--host-reader)", ["--host-reader"])):
            # wall: the process as a user starts it (batches are allocated while the first ones already work); stream: the CLI's own
            # figure with every batch in place before the input is read (--eager-alloc) -- the steady-state rate of a large file
            best, stats, stream = None, [], None
            for eager in (False, False, False, True, True, True):
                t0 = time.time()
                r = subprocess.run([exe, "-b", bam, "-o", out, "-p", "0.8", "-t", str(args.threads), "--stats"] + extra + (["--eager-alloc"] if eager else []),
                                   capture_output=True, text=True)
                dt = time.time() - t0
                assert r.returncode == 0, r.stderr
                lines = r.stderr.strip().splitlines()
                if not eager:
                    best = dt if best is None else min(best, dt)
                    continue
                st = float(lines[-1].split("stream ")[1].split(" s")[0])
                if stream is None or st < stream:
                    stream, stats = st, lines
            got = open(out, "rb").read()
            assert ref_out is None or got == ref_out, "the two readers disagree"
            ref_out = got
            print(f"  {mode}: best of 3 wall {best:.3f} s (default start); steady-state stream (--eager-alloc) {stream:.3f} s -> {hb.n_reads / stream:.3e} alignments/s, {size / stream / 1e9:.2f} GB/s of BAM file")
            for ln in stats[-2:]:
                print(f"    {ln}")
