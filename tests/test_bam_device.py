"""BAM input decoded on the device (exlr_bam_*): BGZF inflate, record walk, SA / CG aux lookup and packing as kernels.
The checker is the batch the BAM was written from -- exactly what the host reader (bam_reader.hpp, tests/test_bam_reader.py)
hands the packer for the same file -- and, end to end, the oracle's lines."""
import struct
import zlib

import numpy as np
import pytest

import oracle_c
from excord_lr_b200 import api, bamio, synth
from excord_lr_b200.batch import ExlrParams, pack_records
from gpu_helpers import gpu_available
from randrec import rand_batch, REF_NAMES

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not gpu_available(), reason="needs a B200")]
FIELDS = ("cigar", "cigar_off", "pos", "tid", "flag", "mapq", "sa_kind", "sa_off", "sa_bytes")


def _header_end(data, blocks):
    """Uncompressed offset of the first record (the host's zlib-side job in the CLI)."""
    u = b"".join(zlib.decompress(data[co:co + cl], -15) for co, cl, _ in blocks[:64])
    l_text, = struct.unpack_from("<i", u, 4)
    n_ref, = struct.unpack_from("<i", u, 8 + l_text)
    o = 12 + l_text
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", u, o)
        o += 8 + l_name
    return o


def _decode(path, params=None, ref_names=REF_NAMES):
    data = open(path, "rb").read()
    blocks, used = api.bgzf_blocks(data)
    p = params or ExlrParams.make()
    ex = api.Extractor(p, ref_names)
    bb = api.BamBatch(ex, len(data) + 64, max(1, len(blocks)))
    bb.load(data[:used], blocks)
    bb.walk(_header_end(data, blocks))
    info = bb.extract()
    return ex, bb, info


def _same(got, hb, n=None):
    n = hb.n_reads if n is None else n
    for f in FIELDS:
        a, b = getattr(got, f), getattr(hb, f)
        if f in ("cigar_off", "sa_off"):
            b = b[:n + 1]
        elif f == "cigar":
            b = b[:int(hb.cigar_off[n])]
        elif f == "sa_bytes":
            b = b[:int(hb.sa_off[n])]
        else:
            b = b[:n]
        assert np.array_equal(a, b), f


@pytest.mark.parametrize("level,block,seq_len", [(1, 0xFF00, 11), (6, 0xFF00, 0), (9, 3000, 37), (0, 0xFF00, 5), (6, 200, 3), (1, 65280, 300)])
def test_decode_matches_the_batch_it_was_written_from(tmp_path, level, block, seq_len):
    # compression levels 0 (stored blocks) .. 9, tiny blocks (fixed Huffman codes, records straddling many blocks), SEQ/QUAL to skip
    hb = synth.with_qnames(synth.config(0, 0.3))
    bam = str(tmp_path / "x.bam")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), level=level, block=block, seq_len=seq_len)
    ex, bb, info = _decode(bam, ref_names=hb.ref_names)
    assert info.status == 0 and info.n_reads == hb.n_reads and info.n_ops == hb.n_ops and info.n_sa_bytes == hb.n_sa_bytes
    assert info.tail_off == info.u_bytes
    got = bb.download(hb.ref_names, hb.n_reads, hb.n_ops, hb.n_sa_bytes)
    _same(got, hb)
    assert got.qnames == hb.qnames
    bb.free(); ex.close()


def test_random_records_and_long_cigars(tmp_path):
    hb = rand_batch(11, 400, qnames=True)
    hb.tid[hb.tid >= len(REF_NAMES)] = 0
    rng = np.random.default_rng(3)
    ops = ((rng.integers(1, 30, 70000).astype(np.uint32) << 4) | rng.choice([0, 1, 2], 70000).astype(np.uint32)).tolist()
    big = pack_records([dict(tid=0, pos=100, flag=0, mapq=60, cigar=ops, sa="chr1,5,+,10M,3,0;"), dict(tid=1, pos=7, flag=16, mapq=1, cigar="5S10M")], REF_NAMES)
    for h, kw in ((hb, dict(seq_len=37, block=4000)), (hb, dict(seq_len=0, level=6)), (big, dict()), (big, dict(seq_len=9, block=1000))):
        bam = str(tmp_path / "r.bam")
        bamio.write_bam(h, bam, **kw)
        ex, bb, info = _decode(bam)
        assert info.status == 0 and info.n_reads == h.n_reads
        _same(bb.download(REF_NAMES, h.n_reads, h.n_ops, h.n_sa_bytes), h)
        bb.free(); ex.close()


@pytest.mark.parametrize("seed", range(30, 42))
def test_random_bams_decode_to_what_was_written(tmp_path, seed):
    rng = np.random.default_rng(seed)
    hb = rand_batch(seed, int(rng.integers(1, 900)), qnames=True)
    hb.tid[hb.tid >= len(REF_NAMES)] = 0
    bam = str(tmp_path / "f.bam")
    bamio.write_bam(hb, bam, level=int(rng.integers(0, 10)), block=int(rng.choice([300, 2000, 20000, 0xFF00])), seq_len=int(rng.choice([0, 1, 33, 700])),
                    random_seq=bool(rng.integers(0, 2)))
    ex, bb, info = _decode(bam)
    assert info.status == 0 and info.n_reads == hb.n_reads
    got = bb.download(REF_NAMES, hb.n_reads, hb.n_ops, hb.n_sa_bytes)
    _same(got, hb)
    assert got.qnames == hb.qnames
    bb.free(); ex.close()


def test_lines_from_compressed_bytes_match_the_oracle(tmp_path):
    # compressed BAM bytes in, the reference's lines out: nothing but block-header hopping happens on the host
    for cfg, scale, seq_len in ((0, 1.0, 20), (1, 0.02, 0), (3, 0.01, 4)):
        hb = synth.config(cfg, scale)
        p = ExlrParams.make(**synth.CONFIGS[cfg]["params"])
        bam = str(tmp_path / "c.bam")
        bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=seq_len, level=6)
        ex, bb, info = _decode(bam, p, hb.ref_names)
        assert info.status == 0 and info.n_reads == hb.n_reads
        res, text = bb.wait_text()
        want = oracle_c.run(hb, p)
        assert res.status == want.status == 0 and res.n_events == len(want.events)
        assert text == oracle_c.format_lines(hb, want.events)
        bb.free(); ex.close()


def test_events_of_a_bam_batch_and_text_beyond_the_first_pinned_buffer(tmp_path):
    # a BAM batch keeps no pinned copies of events / line offsets until exlr_wait asks for them, and its pinned text buffer
    # starts at 16 MB: both paths give what the oracle gives
    hb = synth.config(3, 0.15)                                  # split-heavy: ~0.4 M lines, > 16 MB of text
    p = ExlrParams.make(**synth.CONFIGS[3]["params"])
    bam = str(tmp_path / "d.bam")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=0, level=1)
    data = open(bam, "rb").read()
    blocks, used = api.bgzf_blocks(data)
    ex = api.Extractor(p, hb.ref_names)
    bb = api.BamBatch(ex, len(data) + 64, len(blocks), max_events=2 * hb.n_reads)
    bb.load(data[:used], blocks)
    bb.walk(_header_end(data, blocks))
    info = bb.extract()
    assert info.status == 0 and info.n_reads == hb.n_reads
    want = oracle_c.run(hb, p)
    res, text = bb.wait_text()
    assert res.status == want.status == 0 and res.n_events == len(want.events)
    assert len(text) > (16 << 20)
    assert text == oracle_c.format_lines(hb, want.events)
    r = bb.wait()
    assert r.n_events == len(want.events) and r.events.tobytes() == want.events.tobytes()
    assert np.array_equal(r.line_off, want.line_off)
    bb.free(); ex.close()


def test_partial_record_at_the_end_is_the_tail(tmp_path):
    # a chunk that ends inside a record: every whole record is decoded, tail_off names the first byte of the partial one
    hb = synth.config(0, 0.2)
    bam = str(tmp_path / "t.bam")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), block=5000, seq_len=50)
    data = open(bam, "rb").read()
    blocks, used = api.bgzf_blocks(data)
    keep = len(blocks) * 2 // 3
    ex = api.Extractor(ExlrParams.make(), hb.ref_names)
    bb = api.BamBatch(ex, len(data), len(blocks))
    bb.load(data[:used], blocks[:keep])
    bb.walk(_header_end(data, blocks))
    info = bb.extract()
    assert info.status == 0 and 0 < info.n_reads < hb.n_reads and info.tail_off < info.u_bytes
    n = int(info.n_reads)
    _same(bb.download(hb.ref_names, n, int(info.n_ops), int(info.n_sa_bytes)), hb, n)
    # the next chunk: its own blocks are inflated at once; what the previous walk left over is copied in front on the device
    rest_blocks = blocks[keep:]
    base = rest_blocks[0][0]
    b2 = api.BamBatch(ex, len(data), len(blocks))
    b2.load(data[base:used], [(co - base, cl, ul) for co, cl, ul in rest_blocks])
    b2.walk(0, bb)
    bb, old = b2, bb
    info2 = bb.extract()
    assert info2.status == 0 and info2.n_reads == hb.n_reads - n and info2.tail_off == info2.u_bytes
    rest = bb.download(hb.ref_names, int(info2.n_reads), int(info2.n_ops), int(info2.n_sa_bytes))
    assert np.array_equal(rest.pos, hb.pos[n:]) and np.array_equal(rest.cigar, hb.cigar[int(hb.cigar_off[n]):])
    old.free(); bb.free(); ex.close()


def test_corrupt_block_is_reported(tmp_path):
    hb = synth.config(0, 0.1)
    bam = str(tmp_path / "b.bam")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), block=8000, level=6)
    data = bytearray(open(bam, "rb").read())
    blocks, used = api.bgzf_blocks(bytes(data))
    co, cl, _ = blocks[5]
    for k in range(co + 2, co + min(cl, 40)):
        data[k] ^= 0x5A
    ex = api.Extractor(ExlrParams.make(), hb.ref_names)
    bb = api.BamBatch(ex, len(data) + 65536, len(blocks) + 4)
    bb.load(bytes(data[:used]), blocks)
    bb.walk(0)
    info = bb.extract()
    assert info.status == -7 and info.bad_block == 5
    # a block that inflates fine to the right length but to the wrong bytes: only the CRC-32 tells (htslib checks it too)
    data = bytearray(open(bam, "rb").read())
    raw = zlib.decompress(bytes(data[blocks[7][0]:blocks[7][0] + blocks[7][1]]), -15)
    wrong = bytearray(raw); wrong[100] ^= 1
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp2 = co.compress(bytes(wrong)) + co.flush()
    pre = bytes(data[:blocks[7][0] - 18])
    post = bytes(data[blocks[7][0] + blocks[7][1] + 8:])
    hdr = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp2) + 25)
    forged = pre + hdr + comp2 + struct.pack("<II", zlib.crc32(raw) & 0xFFFFFFFF, len(raw)) + post
    bl2, used2 = api.bgzf_blocks(forged)
    bb.load(forged[:used2], bl2)
    bb.walk(0)
    info = bb.extract()
    assert info.status == -7 and info.bad_block == 7
    ex.set_option(api.EXLR_OPT_BGZF_CRC, 0)
    bb.load(forged[:used2], bl2)
    assert len(bb.inflated()) == sum(b[2] for b in bl2)          # unchecked, the forged block goes through
    bb.free(); ex.close()


@pytest.mark.parametrize("per_chunk,seq_len,block", [(3, 4000, 3000), (1, 4000, 3000), (7, 150, 700), (40, 9000, 0xFF00)])
def test_records_spanning_blocks_and_chunks(tmp_path, per_chunk, seq_len, block):
    # records several BGZF blocks long, chunks of a few blocks: a record can start in one chunk and end two chunks later; the
    # leftover of every walk is handed to the next chunk on the device
    hb = synth.with_qnames(synth.config(0, 0.06))
    bam = str(tmp_path / "s.bam")
    bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=seq_len, block=block, random_seq=True)
    data = open(bam, "rb").read()
    blocks, used = api.bgzf_blocks(data)
    ex = api.Extractor(ExlrParams.make(), hb.ref_names)
    batches = [api.BamBatch(ex, 1 << 22, per_chunk + 1, 0, 1 << 20) for _ in range(3)]
    got_pos, got_cig, got_sa, got_qn, n_chunks = [], [], [], [], 0
    prev, start = None, _header_end(data, blocks)
    for i in range(0, len(blocks), per_chunk):
        bb = batches[n_chunks % 3]
        part = blocks[i:i + per_chunk]
        base = part[0][0]
        end = part[-1][0] + part[-1][1] + 8                 # (the footer belongs to the block: CRC32, ISIZE)
        bb.load(data[base:end], [(co - base, cl, ul) for co, cl, ul in part])
        if prev is None:
            first_u = sum(b[2] for b in blocks[:i])
            bb.walk(max(0, start - first_u))
        else:
            bb.walk(0, prev)
        info = bb.extract()
        assert info.status == 0
        if info.n_reads:
            h = bb.download(hb.ref_names, int(info.n_reads), int(info.n_ops), int(info.n_sa_bytes))
            got_pos.append(h.pos); got_cig.append(h.cigar); got_sa.append(h.sa_bytes); got_qn += h.qnames
        prev, n_chunks = bb, n_chunks + 1
    assert info.tail_off == info.u_bytes
    assert np.array_equal(np.concatenate(got_pos), hb.pos) and np.array_equal(np.concatenate(got_cig), hb.cigar)
    assert np.array_equal(np.concatenate(got_sa), hb.sa_bytes) and got_qn == hb.qnames
    for b in batches:
        b.free()
    ex.close()


def _bgzf(payloads):
    """BGZF blocks around ready-made raw DEFLATE streams: [(deflate bytes, uncompressed bytes)] -> file bytes."""
    out = b""
    for comp, raw in payloads:
        out += (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp +
                struct.pack("<II", zlib.crc32(raw) & 0xFFFFFFFF, len(raw)))
    return out


def test_inflate_against_zlib_on_every_kind_of_deflate_stream():
    # the DEFLATE kernel alone, on streams no BAM writer would bother to produce: every zlib strategy and level, fixed Huffman
    # codes, RLE (distance 1), Huffman-only, stored blocks, sync-flushed streams (empty stored blocks between deflate blocks),
    # maximal matches at maximal distance, incompressible bytes, a single byte, 15-bit codes from a skewed alphabet
    rng = np.random.default_rng(12)
    text = (b"chr2,9877576,+,51421S1538M14S,60,191;" * 700)[:60000]
    skew = bytes(rng.choice(256, 65000, p=np.r_[[0.5, 0.25, 0.125], np.full(253, 0.125 / 253)]).astype(np.uint8))
    far = bytes(rng.integers(0, 256, 300, dtype=np.uint8)) + bytes(32600) + bytes(rng.integers(0, 256, 300, dtype=np.uint8))
    far = (far[:300] + bytes(32468) + far[:300] + far[:258] * 3)[:65536]
    datas = [text, skew, far, bytes(rng.integers(0, 256, 65536, dtype=np.uint8)), b"\x00" * 65536, b"ab" * 32768, b"x",
             bytes(rng.integers(65, 69, 65536, dtype=np.uint8)), bytes(rng.integers(0, 256, 7, dtype=np.uint8)) * 9000]
    payloads = []
    for raw in datas:
        for level, strategy in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (0, zlib.Z_DEFAULT_STRATEGY),
                                (6, zlib.Z_FIXED), (6, zlib.Z_RLE), (6, zlib.Z_HUFFMAN_ONLY), (9, zlib.Z_FILTERED)):
            co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
            payloads.append((co.compress(raw) + co.flush(), raw, f"data{datas.index(raw)} level{level} strategy{strategy}"))
        # several deflate blocks per BGZF block, separated by sync / full flushes (each leaves an empty stored block)
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = b""
        for k in range(0, len(raw), 9000):
            comp += co.compress(raw[k:k + 9000]) + co.flush(zlib.Z_SYNC_FLUSH if (k // 9000) % 2 else zlib.Z_FULL_FLUSH)
        payloads.append((comp + co.flush(), raw, f"data{datas.index(raw)} flushed"))
    payloads = [p for p in payloads if len(p[0]) < 65000]
    data = _bgzf([p[:2] for p in payloads])
    blocks, used = api.bgzf_blocks(data)
    assert used == len(data) and len(blocks) == len(payloads)
    ex = api.Extractor(ExlrParams.make(), REF_NAMES)
    bb = api.BamBatch(ex, len(data) + 64, len(blocks))
    bb.load(data, blocks)
    try:
        got = bb.inflated()
    except api.ExlrError:
        bad = []
        for comp, raw, label in payloads:                   # which ones?
            d1 = _bgzf([(comp, raw)])
            b1, _ = api.bgzf_blocks(d1)
            bb.load(d1, b1)
            try:
                ok = bb.inflated() == raw
            except api.ExlrError:
                ok = False
            if not ok:
                bad.append(label)
        raise AssertionError(f"refused or wrong: {bad}")
    want = b"".join(p[1] for p in payloads)
    assert len(got) == len(want)
    if got != want:
        at, k = 0, 0
        for comp, raw, label in payloads:
            assert got[at:at + len(raw)] == raw, f"block {k} ({label}: {len(raw)} bytes from {len(comp)}) inflated wrong"
            at, k = at + len(raw), k + 1
    # and streams that must be refused: a reserved block type, a stored block whose NLEN does not match, an over-subscribed code,
    # a distance beyond the start of the output, output longer / shorter than ISIZE
    good = zlib.compressobj(6, zlib.DEFLATED, -15)
    gz = good.compress(text) + good.flush()
    bad_streams = [b"\x07" + gz[1:], b"\x01\x05\x00\x00\x00hello", gz[:len(gz) // 2], b"\x03\x02\x00"]
    for i, comp in enumerate(bad_streams):
        d = _bgzf([(comp, text)])
        bl, _ = api.bgzf_blocks(d)
        bb.load(d, bl)
        with pytest.raises(api.ExlrError) as e:
            bb.inflated()
        assert e.value.status == -7, i
    d = _bgzf([(gz, text + b"!")]) + _bgzf([(gz, text[:-1])])
    bl, _ = api.bgzf_blocks(d)
    bb.load(d, bl)
    with pytest.raises(api.ExlrError):
        bb.inflated()
    bb.free(); ex.close()
