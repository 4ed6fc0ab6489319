"""The two CPU restatements against the known-answer vectors (no GPU)."""
import pytest

import oracle as pyoracle
import oracle_c
from excord_lr_b200.batch import ExlrParams, pack_records
from helpers import py_run
from ka_vectors import FFM, KA, REF_NAMES


@pytest.mark.parametrize("cigar,key", FFM)
def test_first_match_pos(cigar, key):
    # reference src/utils.rs:12-42 with the comment example at :50-57
    assert pyoracle.find_first_match_pos(cigar) == key


@pytest.mark.parametrize("ka", KA, ids=[k[0] for k in KA])
def test_ka_c_oracle(ka):
    _, over, rec, want = ka
    hb = pack_records([rec], REF_NAMES)
    p = ExlrParams.make(**over)
    r = oracle_c.run(hb, p)
    assert r.status == 0
    got = oracle_c.format_lines(hb, r.events).decode()
    assert got == "".join(w + "\n" for w in want)
    assert r.line_off.tolist() == [0, len(want)]


@pytest.mark.parametrize("ka", KA, ids=[k[0] for k in KA])
def test_ka_py_oracle(ka):
    _, over, rec, want = ka
    hb = pack_records([rec], REF_NAMES)
    text, err = py_run(hb, ExlrParams.make(**over))
    assert err is None
    assert text == "".join(w + "\n" for w in want)


def test_ka_all_in_one_batch_verbose():
    # all default-parameter vectors in one batch, verbose columns (reference src/utils.rs:205-223, 252-267)
    sel = [k for k in KA if not k[1]]
    hb = pack_records([dict(k[2], qname="read%d" % i) for i, k in enumerate(sel)], REF_NAMES)
    p = ExlrParams.make()
    r = oracle_c.run(hb, p)
    assert r.status == 0
    got = oracle_c.format_lines(hb, r.events, verbose=True).decode()
    text, err = py_run(hb, p, verbose=True)
    assert err is None and got == text
    lines = got.splitlines()
    assert lines[0].endswith("\texcord-lr-alignment-event\tread0\tstrand:1\tflag:0")
    assert any(l.endswith("\texcord-lr-alignment-event\tread1\tstrand:-1\tflag:16") for l in lines)
    assert any("\texcord-lr-split-read\t" in l for l in lines)
    assert any("\texcord-lr-alignment-event-large-ins-one-alignments\t" in l for l in lines)
    assert any("\texcord-lr-alignment-event-large-ins-two-alignments\t" in l for l in lines)
