"""Parity-pinning kit, step 2: the real `excord-lr` binary against the oracle and against the CLI.

Skipped unless oracle/_ref/excord-lr exists (oracle/build_ref.sh builds it where a Rust toolchain is available; this image
has none, so the oracle is pinned only by hand-derived vectors and by its two restatements agreeing -- DESIGN.md 6).  When the
binary IS there, every case writes a synthetic BAM (excord_lr_b200.bamio), runs the reference on it with the config's own
parameters, and requires the oracle's bytes to be identical -- a mismatch is an oracle bug, to be fixed before anything else.
The CLI comparison additionally needs a B200.
"""
import os
import subprocess

import pytest

import oracle_c
from excord_lr_b200 import bamio, synth
from excord_lr_b200.batch import ExlrParams
from gpu_helpers import gpu_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.environ.get("EXLR_REF_BINARY", os.path.join(ROOT, "oracle", "_ref", "excord-lr"))
EXE = os.path.join(ROOT, "excord_lr_b200", "host", "excord-lr-b200")

pytestmark = pytest.mark.skipif(not os.access(REF_EXE, os.X_OK), reason="no reference binary (oracle/build_ref.sh needs cargo): parity unpinned")

# BASELINE.json configs[0..4] at sizes the reference finishes in seconds: (config, scale, reference flags, oracle parameters)
CASES = [
    (0, 1.0, ["-p", "0.8"], dict(max_pct_overlap=0.8)),
    (1, 0.02, ["-p", "0.8"], dict(max_pct_overlap=0.8)),
    (2, 0.002, ["-i", "30", "-n"], dict(indel_min=30)),
    (3, 0.01, ["-s", "-k", "4"], dict(split_only=True, max_supp_alignm=4)),
    (3, 0.01, ["-s", "-k", "8"], dict(split_only=True, max_supp_alignm=8)),
    (4, 0.004, ["-p", "0.8", "-v"], dict(max_pct_overlap=0.8)),
]


def _flags(params):
    return {k: v for k, v in params.items()}


@pytest.mark.parametrize("cfg,scale,flags,params", CASES, ids=[f"c{c[0]}{''.join(c[2])}" for c in CASES])
def test_reference_binary_matches_oracle_and_cli(tmp_path, cfg, scale, flags, params):
    hb = synth.with_qnames(synth.config(cfg, scale))
    bam = str(tmp_path / "in.bam")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=16, level=6)
    verbose = "-v" in flags
    # the reference writes a dbg! line to stderr for every flag-filtered record (src/main.rs:188): discard it
    ref_out = str(tmp_path / "ref.txt")
    r = subprocess.run([REF_EXE, "-b", bam, "-o", ref_out, "-t", "4"] + flags, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    assert r.returncode == 0, "the reference binary failed on a well-formed synthetic BAM"
    ref_bytes = open(ref_out, "rb").read()
    res = oracle_c.run(hb, ExlrParams.make(**params))
    assert res.status == 0
    want = oracle_c.format_lines(hb, res.events, verbose)
    assert want == ref_bytes, "ORACLE BUG: the C restatement differs from the reference binary"
    if gpu_available():
        ours = str(tmp_path / "ours.txt")
        r = subprocess.run([EXE, "-b", bam, "-o", ours] + flags, capture_output=True)
        assert r.returncode == 0, r.stderr
        assert open(ours, "rb").read() == ref_bytes
