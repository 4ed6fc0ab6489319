"""Minimal BGZF/BAM writer (zlib only) used to turn synthetic batches into real BAM files for the reader/packer/CLI
tests, plus the loader of `exlr_bam_dump` output.  Follows SAMv1 §4.2 (BAM) and §4.1 (BGZF); CIGARs with more than
65 535 ops are written with the CG:B,I convention (§4.2.2) that htslib restores transparently."""
from __future__ import annotations

import struct
import zlib
from typing import Optional, Sequence

import numpy as np

from .batch import HostBatch, SA_OTHER, SA_STRING

_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(data: bytes, level: int) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def write_bam(hb: HostBatch, path: str, ref_lens: Optional[Sequence[int]] = None, level: int = 1, block: int = 0xFF00,
              seq_len: int = 0, extra_aux: bytes = b"NM\x43\x05", random_seq: bool = False) -> None:
    """Writes `hb` as a coordinate-order BAM.  seq_len > 0 adds that many bases of SEQ/QUAL per record (the reader must skip them)."""
    names = hb.ref_names
    lens = list(ref_lens) if ref_lens is not None else [2 ** 31 - 1] * len(names)
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in zip(names, lens))
    out = bytearray()
    out += b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(names))
    for n, l in zip(names, lens):
        nb = n.encode() + b"\x00"
        out += struct.pack("<i", len(nb)) + nb + struct.pack("<i", l)
    qn = hb.qnames if hb.qnames is not None else ["r%09d" % i for i in range(hb.n_reads)]
    co, so = hb.cigar_off.tolist(), hb.sa_off.tolist()
    cig = hb.cigar
    sab = hb.sa_bytes.tobytes()
    seq = b"\x12" * ((seq_len + 1) // 2) + b"\x1e" * seq_len
    rng = np.random.default_rng(7)
    with open(path, "wb") as f:
        for i in range(hb.n_reads):
            ops = cig[co[i]:co[i + 1]]
            name = qn[i].encode() + b"\x00"
            aux = bytearray(extra_aux)
            n_ops = len(ops)
            cig_bytes = ops.astype("<u4").tobytes()
            if n_ops > 65535:           # CG convention: placeholder <l_seq>S<reflen>N + the real CIGAR in CG:B,I
                reflen = int(sum(int(v) >> 4 for v in ops if (int(v) & 15) in (0, 2, 3, 7, 8)))
                aux += b"CGBI" + struct.pack("<I", n_ops) + cig_bytes
                cig_bytes = struct.pack("<II", (seq_len << 4) | 4, (min(reflen, (1 << 28) - 1) << 4) | 3)
                n_ops = 2
            if random_seq and seq_len:      # incompressible-ish bases, HiFi-like qualities: realistic BGZF ratios
                seq = (rng.integers(0, 256, (seq_len + 1) // 2, dtype=np.uint8).tobytes() +
                       rng.integers(20, 42, seq_len, dtype=np.uint8).tobytes())
            k = int(hb.sa_kind[i])
            if k == SA_STRING:
                aux += b"SAZ" + sab[so[i]:so[i + 1]] + b"\x00"
            elif k == SA_OTHER:
                aux += b"SAi" + struct.pack("<i", 7)
            body = (struct.pack("<iiBBHHHiiii", int(hb.tid[i]), int(hb.pos[i]), len(name), int(hb.mapq[i]), 4680, n_ops,
                                int(hb.flag[i]), seq_len, -1, -1, 0) + name + cig_bytes + seq + bytes(aux))
            out += struct.pack("<i", len(body)) + body
            while len(out) >= block:
                f.write(_bgzf_block(bytes(out[:block]), level))
                del out[:block]
        if out:
            f.write(_bgzf_block(bytes(out), level))
        f.write(_EOF)


def load_dump(path: str) -> HostBatch:
    """Reads the flat binary file written by host/exlr_bam_dump."""
    with open(path, "rb") as f:
        def arr(dt):
            n = struct.unpack("<Q", f.read(8))[0]
            return np.frombuffer(f.read(n * np.dtype(dt).itemsize), dt).copy()
        names, noff = arr(np.uint8).tobytes(), arr(np.uint32)
        cigar, cigar_off, pos, tid = arr(np.uint32), arr(np.uint64), arr(np.int32), arr(np.int32)
        flag, mapq, sa_kind, sa_off, sa_bytes = arr(np.uint16), arr(np.uint8), arr(np.uint8), arr(np.uint32), arr(np.uint8)
        qn, qoff = arr(np.uint8).tobytes(), arr(np.uint32)
    ref = [names[noff[i]:noff[i + 1]].decode() for i in range(len(noff) - 1)]
    qnames = [qn[qoff[i]:qoff[i + 1]].decode() for i in range(len(qoff) - 1)]
    return HostBatch(cigar, cigar_off, pos, tid, flag, mapq, sa_kind, sa_off, sa_bytes, ref, qnames)


def write_sam(hb: HostBatch, path: str, ref_lens: Optional[Sequence[int]] = None) -> None:
    """The same records as SAM text (what `samtools view -h` would print, SEQ/QUAL '*')."""
    names = hb.ref_names
    lens = list(ref_lens) if ref_lens is not None else [2 ** 31 - 1] * len(names)
    qn = hb.qnames if hb.qnames is not None else ["r%09d" % i for i in range(hb.n_reads)]
    co, so = hb.cigar_off.tolist(), hb.sa_off.tolist()
    sab = hb.sa_bytes.tobytes()
    with open(path, "w") as f:
        f.write("@HD\tVN:1.6\tSO:coordinate\n")
        for n, l in zip(names, lens):
            f.write(f"@SQ\tSN:{n}\tLN:{l}\n")
        for i in range(hb.n_reads):
            ops = hb.cigar[co[i]:co[i + 1]].tolist()
            cig = "".join("%d%s" % (v >> 4, "MIDNSHP=X"[v & 15]) for v in ops) or "*"
            t = int(hb.tid[i])
            cols = [qn[i], str(int(hb.flag[i])), names[t] if t >= 0 else "*", str(int(hb.pos[i]) + 1), str(int(hb.mapq[i])), cig,
                    "*", "0", "0", "*", "*", "NM:i:5"]
            k = int(hb.sa_kind[i])
            if k == SA_STRING:
                cols.append("SA:Z:" + sab[so[i]:so[i + 1]].decode("latin-1"))
            elif k == SA_OTHER:
                cols.append("SA:i:7")
            f.write("\t".join(cols) + "\n")
