"""Committed golden fixtures (tests/golden/, written by tests/golden/make_golden.py): both oracles on CPU, the CUDA path on a GPU.
They pin the restatements against drift; they are NOT reference-generated (the Rust reference cannot be built here)."""
import hashlib
import json
import os

import pytest

import oracle_c
from excord_lr_b200 import synth
from excord_lr_b200.batch import ExlrParams, pack_records
from gpu_helpers import gpu_available
from helpers import py_run
from ka_vectors import REF_NAMES

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def load(name):
    with open(os.path.join(HERE, name)) as f:
        return json.load(f)


def _c1_small():
    g = load("c1_small.json")
    hb = synth.with_qnames(synth.config(0, 0.05))
    assert (hb.n_reads, hb.n_ops) == (g["n_reads"], g["n_ops"]), "the synthetic generator changed: regenerate the fixtures deliberately"
    return g, hb, ExlrParams.make(**g["params"])


def test_ka_vectors_file_both_oracles():
    g = load("ka_vectors.json")
    assert len(g["ka"]) >= 18
    for k in g["ka"]:
        hb = pack_records([k["record"]], REF_NAMES)
        p = ExlrParams.make(**k["params"])
        r = oracle_c.run(hb, p)
        assert r.status == 0
        want = "".join(line + "\n" for line in k["lines"]).encode()
        assert oracle_c.format_lines(hb, r.events) == want, k["id"]
        text, err = py_run(hb, p)
        assert err is None and text.encode() == want, k["id"]


def test_c1_small_oracle_matches_fixture():
    g, hb, p = _c1_small()
    r = oracle_c.run(hb, p)
    assert r.status == 0 and len(r.events) == g["n_lines"]
    assert sha(r.events.tobytes()) == g["events_sha256"] and sha(r.line_off.tobytes()) == g["line_off_sha256"]
    plain = oracle_c.format_lines(hb, r.events, False)
    assert sha(plain) == g["text_sha256"] and plain.decode().splitlines()[:20] == g["first_lines"]
    assert sha(oracle_c.format_lines(hb, r.events, True)) == g["verbose_text_sha256"]
    text, err = py_run(hb, p)
    assert err is None and sha(text.encode()) == g["text_sha256"]


@pytest.mark.gpu
@pytest.mark.skipif(not gpu_available(), reason="needs a B200")
@pytest.mark.parametrize("kernel", [0, 1, 2, 3])
def test_c1_small_cuda_matches_fixture(kernel):
    from excord_lr_b200 import api
    g, hb, p = _c1_small()
    res, text = api.extract(hb, p, 0, kernel, device_format=True)
    assert res.status == 0 and res.n_events == g["n_lines"]
    assert sha(res.events.tobytes()) == g["events_sha256"] and sha(res.line_off.tobytes()) == g["line_off_sha256"]
    assert sha(text) == g["text_sha256"] and sha(res.device_text) == g["text_sha256"]
    _, vtext = api.extract(hb, p, 0, kernel, verbose=True)
    assert sha(vtext) == g["verbose_text_sha256"]


@pytest.mark.gpu
@pytest.mark.skipif(not gpu_available(), reason="needs a B200")
def test_ka_vectors_file_cuda():
    from excord_lr_b200 import api
    for k in load("ka_vectors.json")["ka"]:
        hb = pack_records([k["record"]], REF_NAMES)
        res, text = api.extract(hb, ExlrParams.make(**k["params"]), device_format=True)
        want = "".join(line + "\n" for line in k["lines"]).encode()
        assert res.status == 0 and text == want and res.device_text == want, k["id"]
