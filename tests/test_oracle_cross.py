"""C restatement vs the independent Python restatement on randomized records (no GPU)."""
import pytest

import oracle_c
from excord_lr_b200 import synth
from excord_lr_b200.batch import ExlrParams
from helpers import py_run
from randrec import rand_batch, rand_params

EXLR_ERR_MERGE_DOMAIN = -20


def _compare(hb, p, verbose=False):
    text, err = py_run(hb, p, verbose)
    r = oracle_c.run(hb, p, merge_mode=1)           # literal merge, like the Python restatement
    got = oracle_c.format_lines(hb, r.events, verbose).decode()
    if err is None:
        assert r.status == 0, (r.status, r.err_read)
    else:
        assert r.status < 0 and r.err_read == err
    assert got == text
    # per-record line counts are consistent with the event list
    assert r.line_off[-1] == len(r.events)
    return r


@pytest.mark.parametrize("seed", range(40))
def test_random_default_params(seed):
    _compare(rand_batch(seed, 150), ExlrParams.make(), verbose=(seed % 5 == 0))


@pytest.mark.parametrize("seed", range(40, 100))
def test_random_params(seed):
    _compare(rand_batch(seed, 150, qnames=True), rand_params(seed), verbose=(seed % 7 == 0))


def test_merge_domain_flagged_exactly():
    # merge_mode 0 must flag exactly the records where the literal >2 loop is not the identity
    n_flag = 0
    for seed in range(200, 260):
        hb = rand_batch(seed, 60)
        p = ExlrParams.make(indel_min=1, merge_min=150, exclude_flag=0, mapq=0)
        lit = oracle_c.run(hb, p, merge_mode=1)
        dom = oracle_c.run(hb, p, merge_mode=0)
        if dom.status == EXLR_ERR_MERGE_DOMAIN:
            n_flag += 1
            assert lit.status != 0 and lit.err_read >= dom.err_read or lit.status == 0 or lit.err_read >= dom.err_read
            k = int(dom.err_read)
            assert (dom.events["read_idx"] < k).all()
    assert n_flag > 10


def test_synth_c1_small():
    # BASELINE.json configs[0] generator, reduced: both restatements, default c1 parameters, verbose on
    hb = synth.with_qnames(synth.config(0, 0.05))
    p = ExlrParams.make(**synth.CONFIGS[0]["params"])
    r = _compare(hb, p, verbose=True)
    assert len(r.events) > 50 and r.n_sa_reads > 10


def test_synth_profiles_small():
    for i, scale in ((2, 0.0002), (3, 0.001)):
        hb = synth.config(i, scale)
        r = _compare(hb, ExlrParams.make(**synth.CONFIGS[i]["params"]))
        assert r.n_kept > 0
