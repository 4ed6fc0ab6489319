"""Host-side structure-of-arrays batch of alignment records (the layout of exlr_batch_views,
include/exlr.h) plus the numpy view of the 48-byte exlr_event record.

Field meaning follows the rust_htslib::bam::Record accessors the reference loop reads
(reference src/main.rs:169-212): flags(), mapq(), tid()/contig(), pos(), cigar(), aux(b"SA").
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

EVENT_DTYPE = np.dtype([("lstart", "<i8"), ("lend", "<i8"), ("rstart", "<i8"), ("rend", "<i8"),
                        ("read_idx", "<u4"), ("lchrom", "<u4"), ("rchrom", "<u4"), ("meta", "<u4")])
assert EVENT_DTYPE.itemsize == 48

SA_NONE, SA_STRING, SA_OTHER = 0, 1, 2
BAM_OPS = "MIDNSHP=X"


class ExlrParams(C.Structure):
    """exlr_params (include/exlr.h); defaults = reference Cli defaults (src/main.rs:47-96)."""
    _fields_ = [("mapq", C.c_uint8), ("exclude_secondary", C.c_uint8), ("exclude_unmapped", C.c_uint8),
                ("split_only", C.c_uint8), ("exclude_flag", C.c_uint16), ("reserved0", C.c_uint16),
                ("indel_min", C.c_uint32), ("merge_min", C.c_uint32), ("ins_clip_min", C.c_uint32),
                ("reserved1", C.c_uint32), ("max_pct_overlap", C.c_double), ("max_supp_alignm", C.c_uint64)]

    @classmethod
    def make(cls, mapq=1, exclude_flag=1796, exclude_secondary=False, exclude_unmapped=False, indel_min=50,
             merge_min=5, ins_clip_min=1000, split_only=False, max_pct_overlap=0.0, max_supp_alignm=4):
        return cls(mapq, int(exclude_secondary), int(exclude_unmapped), int(split_only), exclude_flag, 0,
                   indel_min, merge_min, ins_clip_min, 0, float(max_pct_overlap), max_supp_alignm)


assert C.sizeof(ExlrParams) == 40


@dataclass
class HostBatch:
    """Records back to back; cigar_off/sa_off are exclusive prefixes with a trailing total."""
    cigar: np.ndarray       # u32 [n_ops]   BAM encoding len<<4|op
    cigar_off: np.ndarray   # u64 [n+1]
    pos: np.ndarray         # i32 [n]
    tid: np.ndarray         # i32 [n]
    flag: np.ndarray        # u16 [n]
    mapq: np.ndarray        # u8  [n]
    sa_kind: np.ndarray     # u8  [n]
    sa_off: np.ndarray      # u32 [n+1]
    sa_bytes: np.ndarray    # u8  [n_sa]
    ref_names: List[str] = field(default_factory=list)
    qnames: Optional[List[str]] = None

    @property
    def n_reads(self) -> int:
        return int(self.pos.shape[0])

    @property
    def n_ops(self) -> int:
        return int(self.cigar_off[-1])

    @property
    def n_sa_bytes(self) -> int:
        return int(self.sa_off[-1])

    def slice(self, a: int, b: int) -> "HostBatch":
        """Records [a, b) as an independent batch (offsets rebased)."""
        co, so = self.cigar_off, self.sa_off
        return HostBatch(self.cigar[int(co[a]):int(co[b])].copy(), (co[a:b + 1] - co[a]).astype(np.uint64),
                         self.pos[a:b].copy(), self.tid[a:b].copy(), self.flag[a:b].copy(), self.mapq[a:b].copy(),
                         self.sa_kind[a:b].copy(), (so[a:b + 1] - so[a]).astype(np.uint32),
                         self.sa_bytes[int(so[a]):int(so[b])].copy(), self.ref_names,
                         None if self.qnames is None else self.qnames[a:b])

    @staticmethod
    def concat(parts: Sequence["HostBatch"]) -> "HostBatch":
        """Batches back to back (offsets rebased); used to build one rank's shard from its round-robin batches."""
        if not parts:
            raise ValueError("nothing to concatenate")
        co, so, cbase, sbase = [np.zeros(1, np.uint64)], [np.zeros(1, np.uint32)], 0, 0
        for p in parts:
            co.append(p.cigar_off[1:] + np.uint64(cbase))
            so.append((p.sa_off[1:].astype(np.uint64) + sbase).astype(np.uint32))
            cbase += p.n_ops
            sbase += p.n_sa_bytes
        cat = lambda f: np.concatenate([getattr(p, f) for p in parts])
        qn = None if any(p.qnames is None for p in parts) else [q for p in parts for q in p.qnames]
        return HostBatch(cat("cigar"), np.concatenate(co), cat("pos"), cat("tid"), cat("flag"), cat("mapq"), cat("sa_kind"),
                         np.concatenate(so), cat("sa_bytes"), parts[0].ref_names, qn)

    def qname_blob(self):
        """(bytes, u32 offsets[n+1]) for the verbose formatter."""
        names = self.qnames if self.qnames is not None else ["r%09d" % i for i in range(self.n_reads)]
        off = np.zeros(self.n_reads + 1, np.uint32)
        enc = [s.encode() for s in names]
        off[1:] = np.cumsum([len(e) for e in enc], dtype=np.uint64).astype(np.uint32)
        return b"".join(enc), off


def parse_cigar_string(s: str) -> List[int]:
    """'100M60D' -> BAM-encoded u32 ops."""
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | BAM_OPS.index(ch))
            num = ""
    return out


def pack_records(records: Sequence[dict], ref_names: Sequence[str]) -> HostBatch:
    """records: dicts with contig (name or None), pos, flag, mapq, cigar (str or list of u32),
    sa (str|None), sa_kind (optional), qname (optional)."""
    name_to_tid = {n: i for i, n in enumerate(ref_names)}
    cig, coff, pos, tid, flag, mapq, sk, soff, sab, qn = [], [0], [], [], [], [], [], [0], bytearray(), []
    for i, r in enumerate(records):
        ops = parse_cigar_string(r["cigar"]) if isinstance(r["cigar"], str) else list(r["cigar"])
        cig.extend(ops)
        coff.append(len(cig))
        pos.append(r["pos"])
        c = r.get("contig")
        tid.append(r["tid"] if "tid" in r else (-1 if c is None else name_to_tid[c]))
        flag.append(r["flag"])
        mapq.append(r["mapq"])
        sa = r.get("sa")
        kind = r.get("sa_kind", SA_NONE if sa is None else SA_STRING)
        sk.append(kind)
        if sa is not None and kind == SA_STRING:
            sab.extend(sa.encode("latin-1"))
        soff.append(len(sab))
        qn.append(r.get("qname", "r%09d" % i))
    return HostBatch(np.array(cig, np.uint32), np.array(coff, np.uint64), np.array(pos, np.int32),
                     np.array(tid, np.int32), np.array(flag, np.uint16), np.array(mapq, np.uint8),
                     np.array(sk, np.uint8), np.array(soff, np.uint32),
                     np.frombuffer(bytes(sab), np.uint8).copy(), list(ref_names), qn)
