"""Reader + packer (host C++, no GPU): synthetic batch -> BAM (Python writer) -> exlr_bam_dump -> identical batch."""
import os
import subprocess

import numpy as np
import pytest

from excord_lr_b200 import bamio, synth
from excord_lr_b200.batch import pack_records
from randrec import rand_batch, REF_NAMES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMP = os.path.join(ROOT, "excord_lr_b200", "host", "exlr_bam_dump")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "excord_lr_b200", "host"), "exlr_bam_dump"], stdout=subprocess.DEVNULL)


def _roundtrip(hb, tmp_path, threads=4, **kw):
    _build()
    bam, out = str(tmp_path / "x.bam"), str(tmp_path / "x.bin")
    bamio.write_bam(hb, bam, **kw)
    subprocess.check_call([DUMP, bam, out, str(threads)])
    got = bamio.load_dump(out)
    for f in ("cigar", "cigar_off", "pos", "tid", "flag", "mapq", "sa_kind", "sa_off", "sa_bytes"):
        assert np.array_equal(getattr(got, f), getattr(hb, f)), f
    assert got.ref_names == hb.ref_names
    return got


def test_roundtrip_config0(tmp_path):
    hb = synth.with_qnames(synth.config(0, 0.2))
    got = _roundtrip(hb, tmp_path, ref_lens=synth.ref_lens())
    assert got.qnames == hb.qnames


def test_roundtrip_random_records_with_seq_and_threads(tmp_path):
    hb = rand_batch(11, 400, qnames=True)
    hb.tid[hb.tid >= len(REF_NAMES)] = 0
    _roundtrip(hb, tmp_path, threads=1, seq_len=37, block=4000)      # small blocks: records straddle BGZF blocks
    _roundtrip(hb, tmp_path, threads=8, seq_len=0, level=6)


def test_long_cigar_cg_tag_restore(tmp_path):
    # > 65 535 ops: written as <l_seq>S<reflen>N + CG:B,I, restored by the reader like htslib's bam_read1
    rng = np.random.default_rng(3)
    ops = ((rng.integers(1, 30, 70000).astype(np.uint32) << 4) | rng.choice([0, 1, 2], 70000).astype(np.uint32)).tolist()
    hb = pack_records([dict(tid=0, pos=100, flag=0, mapq=60, cigar=ops, sa="chr1,5,+,10M,3,0;"),
                       dict(tid=1, pos=7, flag=16, mapq=1, cigar="5S10M")], REF_NAMES)
    got = _roundtrip(hb, tmp_path)
    assert int(got.cigar_off[1]) == 70000


def test_truncated_bam_stops_quietly(tmp_path):
    # the reference breaks out of its loop on a read error and exits 0 with what it had (src/main.rs:165-168)
    _build()
    hb = synth.config(0, 0.2)
    bam = str(tmp_path / "t.bam")
    bamio.write_bam(hb, bam, block=8000)
    data = open(bam, "rb").read()
    open(bam, "wb").write(data[: len(data) * 2 // 3])
    rc = subprocess.call([DUMP, bam, str(tmp_path / "t.bin"), "2"], stderr=subprocess.DEVNULL)
    assert rc == 4
    got = bamio.load_dump(str(tmp_path / "t.bin"))
    n = got.n_reads
    assert 0 < n < hb.n_reads and np.array_equal(got.pos, hb.pos[:n])


def test_sam_text_input(tmp_path):
    # SAM text is accepted like BAM (the reference reads .sam through htslib's auto-detection, reference Makefile:73-74)
    _build()
    hb = synth.with_qnames(synth.config(0, 0.1))
    hb2 = rand_batch(5, 200, qnames=True)
    hb2.tid[hb2.tid >= len(REF_NAMES)] = 0
    for k, h in enumerate((hb, hb2)):
        valid = (h.cigar & 15) <= 8
        assert valid.all()
        sam, out = str(tmp_path / f"x{k}.sam"), str(tmp_path / f"x{k}.bin")
        bamio.write_sam(h, sam)
        subprocess.check_call([DUMP, sam, out, "2"])
        got = bamio.load_dump(out)
        for f in ("cigar", "cigar_off", "pos", "tid", "flag", "mapq", "sa_kind", "sa_off", "sa_bytes"):
            assert np.array_equal(getattr(got, f), getattr(h, f)), f
        assert got.qnames == h.qnames and got.ref_names == h.ref_names


def test_stream_from_stdin(tmp_path):
    # "-" reads a pipe: the format sniff pushes its bytes back instead of seeking (BAM and SAM alike)
    _build()
    hb = synth.with_qnames(synth.config(0, 0.1))
    bam, sam = str(tmp_path / "s.bam"), str(tmp_path / "s.sam")
    bamio.write_bam(hb, bam, seq_len=9)
    bamio.write_sam(hb, sam)
    for src in (bam, sam):
        out = str(tmp_path / "s.bin")
        with open(src, "rb") as f:
            subprocess.check_call([DUMP, "-", out, "3"], stdin=f)
        got = bamio.load_dump(out)
        for f in ("cigar", "cigar_off", "pos", "tid", "flag", "mapq", "sa_kind", "sa_off", "sa_bytes"):
            assert np.array_equal(getattr(got, f), getattr(hb, f)), f


def test_cg_tag_guards_like_htslib(tmp_path):
    # bam_tag2cigar (inside htslib's bam_read1) swaps the CIGAR in only if CG:B,I holds at least n_cigar ops (and < 2^29):
    # a <l_seq>S first op with a shorter CG array keeps the CIGAR as stored
    import struct
    hb = pack_records([dict(tid=0, pos=100, flag=0, mapq=60, cigar="5S10M"), dict(tid=1, pos=7, flag=16, mapq=1, cigar="5S3M")], REF_NAMES)
    short_cg = b"CGBI" + struct.pack("<II", 1, (7 << 4) | 0)
    got = _roundtrip(hb, tmp_path, seq_len=5, extra_aux=short_cg)
    assert got.cigar.tolist() == hb.cigar.tolist()
    # ... and with enough ops it is swapped in (2 >= n_cigar = 2)
    long_cg = b"CGBI" + struct.pack("<III", 2, (7 << 4) | 0, (9 << 4) | 2)
    bam, out = str(tmp_path / "cg.bam"), str(tmp_path / "cg.bin")
    bamio.write_bam(hb, bam, seq_len=5, extra_aux=long_cg)
    subprocess.check_call([DUMP, bam, out, "1"])
    got = bamio.load_dump(out)
    assert got.cigar.tolist() == [(7 << 4) | 0, (9 << 4) | 2] * 2


def test_chunker_sees_every_block(tmp_path):
    # the host side of the GPU BAM decoder: whatever the chunk geometry, the same blocks in the same order (checked against
    # the Python header hop of excord_lr_b200.api.bgzf_blocks), and the BAM header parsed with zlib
    import struct
    import zlib
    from excord_lr_b200 import api
    _build()
    hb = synth.with_qnames(synth.config(0, 0.3))
    bam = str(tmp_path / "k.bam")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=40, block=3000)
    data = open(bam, "rb").read()
    blocks, used = api.bgzf_blocks(data)
    assert used == len(data)
    u = b"".join(zlib.decompress(data[co:co + cl], -15) for co, cl, _ in blocks[:40])
    l_text, = struct.unpack_from("<i", u, 4)
    o = 12 + l_text
    for _ in range(len(hb.ref_names)):
        o += 8 + struct.unpack_from("<i", u, o)[0]
    uoff = np.concatenate([[0], np.cumsum([b[2] for b in blocks])])
    first = int(np.searchsorted(uoff, o, side="right") - 1)
    outs = set()
    for cap, capb, th in ((8 << 20, 4096, 4), (70000, 4096, 1), (1 << 20, 7, 3), (200000, 2, 2)):
        r = subprocess.run([DUMP, "--chunker", bam, str(cap), str(capb), str(th)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        l1, l2 = r.stdout.strip().splitlines()
        assert l1 == f"refs {len(hb.ref_names)} first_record_off {o - uoff[first]}"
        outs.add(l2)
    want_blocks = blocks[first:]
    assert len(outs) == 1
    assert outs.pop().startswith(f"blocks {len(want_blocks)} ulen {sum(b[2] for b in want_blocks)} clen {sum(b[1] for b in want_blocks)} ")
    sam = str(tmp_path / "k.sam")
    bamio.write_sam(hb, sam)
    assert subprocess.run([DUMP, "--chunker", sam], capture_output=True, text=True).returncode == 1


def test_chunker_reader_pool_reads_the_same_bytes(tmp_path):
    # chunks larger than one read slice are read by the stream's worker pool (slices taken by whoever is free, the same workers
    # chunk after chunk): any thread count and chunk size sees the same blocks, and the sample hash covers their bytes
    _build()
    hb = synth.config(1, 0.001)
    bam = str(tmp_path / "p.bam")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=15000, random_seq=True, level=1)
    assert os.path.getsize(bam) > (12 << 20)
    outs = set()
    for cap, th in ((5 << 20, 1), (5 << 20, 8), (9 << 20, 32), (64 << 20, 5), (3 << 20, 16)):
        r = subprocess.run([DUMP, "--chunker", bam, str(cap), "30000", str(th)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.add(r.stdout)
    assert len(outs) == 1
