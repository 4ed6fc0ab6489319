"""The CLI front end (host C++).  Startup behaviour is checked without a GPU; the full BAM -> event file runs are GPU tests
and must reproduce the oracle's bytes exactly."""
import os
import subprocess

import pytest

import oracle_c
from excord_lr_b200 import bamio, synth
from excord_lr_b200.batch import ExlrParams, pack_records
from gpu_helpers import gpu_available
from randrec import REF_NAMES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "excord_lr_b200", "host", "excord-lr-b200")


@pytest.fixture(scope="module", autouse=True)
def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "excord_lr_b200", "host")], stdout=subprocess.DEVNULL)


def run(args, **kw):
    return subprocess.run([EXE] + args, capture_output=True, text=True, **kw)


def test_startup_messages_and_exit_codes(tmp_path):
    # reference src/main.rs:113-134: messages go to stdout, exit code 1
    r = run(["-b", str(tmp_path / "nope.bam"), "-o", str(tmp_path / "o.txt")])
    assert r.returncode == 1 and r.stdout == f"Ivalid BAM file path: {tmp_path / 'nope.bam'} \n"
    bam = tmp_path / "x.bam"
    bam.write_bytes(b"")
    r = run(["-b", str(bam), "-o", str(tmp_path / "missing_dir" / "o.txt")])
    assert r.returncode == 1 and r.stdout == f"Output directory does not exists: {tmp_path / 'missing_dir'} \n"
    cram = tmp_path / "x.cram"
    cram.write_bytes(b"")
    r = run(["-b", str(cram), "-o", str(tmp_path / "o.txt")])
    assert r.returncode == 1 and r.stdout == "excord-lr is running on CRAM file, reference(-r) is required.\n"
    assert (tmp_path / "o.txt").exists()            # the output file is created before the input is opened (main.rs:135)
    assert run(["-V"]).stdout == "excord-LR 0.1.17\n"
    assert run(["--help"]).returncode == 0
    assert run(["-o", "x"]).returncode == 2 and run(["-b", "x", "-o", "y", "--bogus"]).returncode == 2


def test_debug_prints_parsed_cli(tmp_path):
    r = run(["-d", "-b", str(tmp_path / "nope.bam"), "-o", str(tmp_path / "o.txt"), "-Q", "3", "-i", "30", "-n", "-p", "0.8", "-k8"])
    assert r.stdout.startswith('Cli { bam: "') and "mapq: 3" in r.stdout and "indel_min: 30" in r.stdout
    assert "not_merge: true" in r.stdout and "max_pct_overlap: 0.8" in r.stdout and "max_supp_alignm: 8" in r.stdout


gpu = [pytest.mark.gpu, pytest.mark.skipif(not gpu_available(), reason="needs a B200")]


def _expect(hb, p, verbose=False):
    r = oracle_c.run(hb, p)
    return r, oracle_c.format_lines(hb, r.events, verbose)      # (on a panic the oracle keeps exactly the lines written by then)


@pytest.mark.gpu
@pytest.mark.skipif(not gpu_available(), reason="needs a B200")
@pytest.mark.parametrize("extra,params,verbose", [
    (["-p", "0.8"], dict(max_pct_overlap=0.8), False),
    (["-p", "0.8", "-v", "--batch-reads", "777"], dict(max_pct_overlap=0.8), True),
    (["--pct-overlap=0.8", "-s", "-k", "8", "-t", "2", "--batch-reads", "100"], dict(max_pct_overlap=0.8, split_only=True, max_supp_alignm=8), False),
    (["-i", "30", "-n", "-Q", "0", "-F", "0", "-S", "-U", "--ins-clip-min", "40", "--verbose"], dict(indel_min=30, mapq=0, exclude_flag=0, exclude_secondary=True, exclude_unmapped=True, ins_clip_min=40), True),
])
def test_bam_to_event_file_matches_oracle(tmp_path, extra, params, verbose):
    hb = synth.with_qnames(synth.config(0, 1.0))
    if params.get("exclude_flag") == 0:
        hb.tid[hb.tid < 0] = 0                      # with -F 0 an unplaced record would reach contig() and panic
    bam, out = str(tmp_path / "c1.bam"), str(tmp_path / "c1.txt")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=11)
    r = run(["-b", bam, "-o", out] + extra)
    assert r.returncode == 0, r.stderr
    want_r, want = _expect(hb, ExlrParams.make(**params), verbose)
    assert want_r.status == 0 and len(want) > 1000
    assert open(out, "rb").read() == want


@pytest.mark.gpu
@pytest.mark.skipif(not gpu_available(), reason="needs a B200")
def test_reference_panic_gives_partial_output_and_exit_101(tmp_path):
    recs = [dict(tid=0, pos=100 + i, flag=0, mapq=60, cigar="100M60D200M", qname="a%d" % i) for i in range(500)]
    recs[321] = dict(tid=0, pos=5, flag=0, mapq=60, cigar="10M", sa="chr1,5,*,10M,3,0;", qname="bad")
    hb = pack_records(recs, REF_NAMES)
    bam, out = str(tmp_path / "p.bam"), str(tmp_path / "p.txt")
    bamio.write_bam(hb, bam)
    r = run(["-b", bam, "-o", out, "--batch-reads", "64"])
    assert r.returncode == 101 and "strand" in r.stderr
    want_r, want = _expect(hb, ExlrParams.make())
    assert want_r.status == -14 and want_r.err_read == 321
    assert open(out, "rb").read() == want


def _n_devices():
    try:
        from excord_lr_b200 import api
        return api.load_library().exlr_device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(not gpu_available(), reason="needs a B200")
def test_sam_text_to_event_file(tmp_path):
    # the reference's Makefile feeds it a .sam (Makefile:73-74): same bytes out as for the BAM
    hb = synth.with_qnames(synth.config(0, 0.3))
    sam, out = str(tmp_path / "c1.sam"), str(tmp_path / "c1.txt")
    bamio.write_sam(hb, sam, ref_lens=synth.ref_lens())
    for extra, params, verbose in ((["-p", "0.8"], dict(max_pct_overlap=0.8), False), (["-v", "-i", "30"], dict(indel_min=30), True)):
        r = run(["-b", sam, "-o", out] + extra)
        assert r.returncode == 0, r.stderr
        want_r, want = _expect(hb, ExlrParams.make(**params), verbose)
        assert want_r.status == 0 and len(want) > 500
        assert open(out, "rb").read() == want


@pytest.mark.gpu
@pytest.mark.skipif(not gpu_available(), reason="needs a B200")
def test_stdin_to_stdout_streaming(tmp_path):
    # `-b -` / `-o -` (beyond the reference, which insists on regular files): BAM in through a pipe, lines out through a pipe
    hb = synth.with_qnames(synth.config(0, 0.3))
    bam = str(tmp_path / "c1.bam")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=7)
    want_r, want = _expect(hb, ExlrParams.make(max_pct_overlap=0.8))
    with open(bam, "rb") as f:
        r = subprocess.run([EXE, "-b", "-", "-o", "-", "-p", "0.8", "--batch-reads", "500"], stdin=f, capture_output=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == want
    with open(bam, "rb") as f:
        r = subprocess.run([EXE, "-b", "/dev/stdin", "-o", str(tmp_path / "o.txt"), "-p", "0.8"], stdin=f, capture_output=True)
    assert r.returncode == 0 and open(tmp_path / "o.txt", "rb").read() == want
    # a path that does not exist still gets the reference's message
    r = run(["-b", str(tmp_path / "nope.bam"), "-o", "-"])
    assert r.returncode == 1 and r.stdout.startswith("Ivalid BAM file path")


@pytest.mark.gpu
@pytest.mark.skipif(not gpu_available(), reason="needs a B200")
def test_event_buffers_grow_instead_of_failing(tmp_path):
    # an event-dense input overflows the batch's event buffers: the CLI grows them and runs the batch again (the reference has
    # no such limit, so the run must complete with identical bytes)
    hb = synth.with_qnames(synth.config(0, 0.5))
    bam, out = str(tmp_path / "c1.bam"), str(tmp_path / "c1.txt")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens())
    for extra, params, verbose in ((["-i", "1", "--batch-events", "64", "--stats"], dict(indel_min=1), False),
                                   (["-i", "1", "-v", "--batch-events", "64", "--batch-reads", "999", "--stats"], dict(indel_min=1), True),
                                   (["-i", "2", "-m", "5", "--batch-events", "100", "--stats"], dict(indel_min=2, merge_min=5), False)):
        r = run(["-b", bam, "-o", out] + extra)
        assert r.returncode == 0, r.stderr
        assert "re-run with larger event buffers" in r.stderr and " (0 re-run" not in r.stderr
        want_r, want = _expect(hb, ExlrParams.make(**params), verbose)
        assert want_r.status == 0 and len(want) > 100000
        assert open(out, "rb").read() == want


@pytest.mark.gpu
@pytest.mark.skipif(_n_devices() < 2, reason="needs two B200s")
def test_two_gpus_round_robin_same_bytes(tmp_path):
    # in-process sharding: batch k on GPU k % 2, re-assembled in order by the writer thread
    hb = synth.with_qnames(synth.config(0, 1.0))
    bam, out = str(tmp_path / "c1.bam"), str(tmp_path / "c1.txt")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens())
    want_r, want = _expect(hb, ExlrParams.make(max_pct_overlap=0.8))
    for extra in (["--gpus", "2", "--batch-reads", "700"], ["--gpus", "2", "--batch-reads", "700", "-t", "2"], ["--batch-reads", "333"]):
        r = run(["-b", bam, "-o", out, "-p", "0.8", "--stats"] + extra)
        assert r.returncode == 0, r.stderr
        assert open(out, "rb").read() == want
        assert "on 2 of" in r.stderr or "--gpus" not in extra
    want_r, want = _expect(hb, ExlrParams.make(max_pct_overlap=0.8), True)
    r = run(["-b", bam, "-o", out, "-p", "0.8", "-v", "--gpus", "2", "--batch-reads", "1000"])
    assert r.returncode == 0 and open(out, "rb").read() == want


@pytest.mark.gpu
@pytest.mark.skipif(not gpu_available(), reason="needs a B200")
def test_gpu_bam_decoder_chunks_and_host_reader_agree(tmp_path):
    # BGZF BAM in a regular file is inflated and walked on the GPU (exlr_bam_*); --host-reader keeps it on zlib threads.
    # Small chunks make records straddle chunk boundaries (the tail record's blocks are repeated in front of the next chunk).
    hb = synth.with_qnames(synth.config(0, 1.0))
    bam, out = str(tmp_path / "c1.bam"), str(tmp_path / "c1.txt")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=120, block=6000, level=6)
    for verbose in (False, True):
        want_r, want = _expect(hb, ExlrParams.make(max_pct_overlap=0.8), verbose)
        assert want_r.status == 0
        for extra in (["--chunk-blocks", "9"], ["--chunk-blocks", "64", "--gpus", "1"], ["--chunk-mb", "1"], [], ["--host-reader", "-t", "3"]):
            r = run(["-b", bam, "-o", out, "-p", "0.8", "--stats"] + (["-v"] if verbose else []) + extra)
            assert r.returncode == 0, r.stderr
            assert open(out, "rb").read() == want, extra
            assert ("GPU BAM decoder" in r.stderr) == ("--host-reader" not in extra)


@pytest.mark.gpu
@pytest.mark.skipif(not gpu_available(), reason="needs a B200")
def test_gpu_bam_decoder_damaged_files(tmp_path):
    # truncated file / corrupt block: the reference's reader returns an error and the loop ends quietly with the lines so far
    # (src/main.rs:165-168); both readers must stop at the same record
    hb = synth.with_qnames(synth.config(0, 0.5))
    bam, out = str(tmp_path / "d.bam"), str(tmp_path / "d.txt")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), block=8000)
    data = open(bam, "rb").read()
    trunc = str(tmp_path / "t.bam")
    open(trunc, "wb").write(data[: len(data) * 2 // 3])
    outs = []
    for extra in (["--host-reader"], ["--chunk-blocks", "16"], []):
        r = run(["-b", trunc, "-o", out] + extra)
        assert r.returncode == 0, r.stderr
        outs.append(open(out, "rb").read())
    assert len(outs[0]) > 1000 and outs[0] == outs[1] == outs[2]
    full_r, full = _expect(hb, ExlrParams.make())
    assert full.startswith(outs[0]) and len(outs[0]) < len(full)
    # event-dense chunk: the event buffers of a BAM chunk grow too
    r = run(["-b", bam, "-o", out, "-i", "1", "--batch-events", "64", "--stats"])
    want_r, want = _expect(hb, ExlrParams.make(indel_min=1))
    assert r.returncode == 0 and open(out, "rb").read() == want and " (0 re-run" not in r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not gpu_available(), reason="needs a B200")
def test_gpu_bam_decoder_odd_files(tmp_path):
    # header-only BAM, a header larger than a BGZF block, records several blocks long (ONT-sized SEQ/QUAL), unplaced records
    from excord_lr_b200.batch import HostBatch
    import numpy as np
    out = str(tmp_path / "o.txt")
    # 1. no records at all
    hb = synth.config(0, 0.01)
    empty = hb.slice(0, 0)
    bam = str(tmp_path / "empty.bam")
    bamio.write_bam(empty, bam, ref_lens=synth.ref_lens())
    for extra in ([], ["--host-reader"]):
        r = run(["-b", bam, "-o", out] + extra)
        assert r.returncode == 0, r.stderr
        assert open(out, "rb").read() == b""
    # 2. 3 000 contigs: the header spans several BGZF blocks and ends inside one
    names = ["chr%d_random_contig_with_a_long_name_%d" % (i, i * 7919) for i in range(3000)]
    recs = [dict(tid=i * 37 % 3000, pos=1000 + i, flag=0, mapq=60, cigar="100M60D200M%dS" % (2000 + i),
                 sa="%s,%d,+,2000S300M,60,1;" % (names[(i * 11) % 3000], 5000 + i), qname="q%d" % i) for i in range(400)]
    big = pack_records(recs, names)
    bam = str(tmp_path / "bighdr.bam")
    bamio.write_bam(big, bam, seq_len=33)
    want_r, want = _expect(big, ExlrParams.make())
    assert want_r.status == 0 and len(want) > 10000
    for extra in ([], ["--chunk-blocks", "5"], ["--host-reader"]):
        r = run(["-b", bam, "-o", out] + extra)
        assert r.returncode == 0, r.stderr
        assert open(out, "rb").read() == want, extra
    # 3. records of ~300 KB (five BGZF blocks each), chunks smaller than a record's span
    hb = synth.with_qnames(synth.config(0, 0.01))
    bam = str(tmp_path / "long.bam")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=200_000, random_seq=True)
    want_r, want = _expect(hb, ExlrParams.make(max_pct_overlap=0.8), True)
    for extra in ([], ["--chunk-blocks", "3"], ["--chunk-blocks", "16", "--max-record-mb", "1"]):
        r = run(["-b", bam, "-o", out, "-p", "0.8", "-v"] + extra)
        assert r.returncode == 0, r.stderr
        assert open(out, "rb").read() == want, extra
