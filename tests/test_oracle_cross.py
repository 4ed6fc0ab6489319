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
    r = oracle_c.run(hb, p)
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


def test_literal_merge_loop_beyond_the_identity():
    # merge_min > 2 * indel_min: the >2 loop (main.rs:636-742) merges, duplicates or panics; both restatements run it literally
    n_panic = n_changed = 0
    for seed in range(200, 260):
        hb = rand_batch(seed, 60)
        p = ExlrParams.make(indel_min=1, merge_min=150, exclude_flag=0, mapq=0)
        r = _compare(hb, p)
        n_panic += r.status == EXLR_ERR_MERGE_DOMAIN
        n_changed += r.status == 0
    assert n_panic > 5 and n_changed >= 1
    # four close Dels -> [ab, bc, bc, cd]: duplicates, no panic
    from excord_lr_b200.batch import pack_records
    from randrec import REF_NAMES
    hb = pack_records([dict(tid=0, pos=1000, flag=0, mapq=60, cigar="100M60D2M60D2M60D2M60D100M")], REF_NAMES)
    r = _compare(hb, ExlrParams.make(merge_min=200))
    assert r.status == 0 and len(r.events) == 4
    assert r.events["lend"].tolist() == [1100, 1162, 1162, 1224] and r.events["rstart"].tolist() == [1222, 1284, 1284, 1346]


def test_indel_arm_panic_keeps_the_sa_arm_lines():
    # the SA arm's f.write calls (main.rs:395-515) precede the indel arm (main.rs:523-742): when the merge loop panics the
    # record's split / large-INS lines are already in the BufWriter and reach the file on unwind
    from excord_lr_b200.batch import pack_records
    from randrec import REF_NAMES
    recs = [dict(tid=0, pos=100, flag=0, mapq=60, cigar="100M60D200M"),
            dict(tid=1, pos=5000, flag=0, mapq=60, cigar="10M60D2M60D2M60D10M3000S", sa="chr2,7001,+,3000S500M,60,1;"),
            dict(tid=0, pos=900, flag=0, mapq=60, cigar="100M60D200M")]
    hb = pack_records(recs, REF_NAMES)
    r = _compare(hb, ExlrParams.make(merge_min=200))
    assert r.status == EXLR_ERR_MERGE_DOMAIN and r.err_read == 1 and r.n_err_lines == 1 and len(r.events) == 2
    assert r.line_off.tolist() == [0, 1, 2, 2]


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
