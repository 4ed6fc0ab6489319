"""oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Second, independent CPU restatement of excord-lr's per-record loop body, written straight
from the Rust (reference src/main.rs:158-770, src/utils.rs, src/aligments_event.rs,
src/split_read_event.rs) and working on *text* the way the reference does (CIGAR strings,
HashMap<char,u32> buckets, String chrom names), whereas oracle/exlr_oracle.c works on the
packed structure-of-arrays batch and emits binary event records.  The two are compared on
randomized records in tests/; agreement of two independently written restatements plus the
hand-derived known answers is what pins parity, because the
reference has no tests/fixtures and cannot be built here (PARITY UNPINNED).

Only tests/ may import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

U32 = 0xFFFFFFFF
BAM_OPS = "MIDNSHP=X"          # rust-htslib Cigar enum order (main.rs:226-234)
OP_CHARS = set("DMIHSPX=N")    # utils.rs:16-24


class ReferencePanic(Exception):
    """The reference would panic (exit 101) on this record."""

    def __init__(self, what: str, read: int = -1):
        super().__init__(what)
        self.what = what
        self.read = read


@dataclass
class Params:                   # Cli, main.rs:37-105
    mapq: int = 1
    exclude_flag: int = 1796
    exclude_secondary: bool = False
    exclude_unmapped: bool = False
    indel_min: int = 50
    merge_min: int = 5
    ins_clip_min: int = 1000
    split_only: bool = False
    max_pct_overlap: float = 0.0
    max_supp_alignm: int = 4
    verbose: bool = False


@dataclass
class Record:                   # what the loop reads off rust_htslib::bam::Record
    contig: Optional[str]       # None <=> tid < 0
    pos: int                    # 0-based
    flag: int
    mapq: int
    cigar: List[Tuple[int, int]]    # (bam op code, len)
    sa: Optional[str] = None        # SA:Z value, None if no SA aux
    sa_is_string: bool = True       # False: SA aux of a non-Z type (main.rs:308 `if let` fails)
    qname: str = "q"


@dataclass
class Seg:                      # SplitReadEvent, split_read_event.rs:3-11
    chrom: str
    start: int
    end: int
    cigar_map: dict
    strand: int
    mapq: int
    raw_cigar: str


def _strip(chrom: str) -> str:  # aligments_event.rs:38-42
    return chrom[3:] if chrom.startswith("chr") else chrom


def _new_seg(chrom, start, cigar_map, strand, mapq, cigar_string) -> Seg:   # split_read_event.rs:14-45
    end = start + cigar_map["D"] + cigar_map["M"] + cigar_map["="] + cigar_map["X"] + -1
    return Seg(_strip(chrom), start, end + 1, cigar_map, strand, mapq, cigar_string)


def _rust_parse_int(s: str, lo: int, hi: int, signed: bool) -> int:
    """str::parse::<iN/uN>: optional '+' (and '-' for signed), then ASCII digits only."""
    t = s
    neg = False
    if t[:1] == "+":
        t = t[1:]
    elif t[:1] == "-" and signed:
        neg = True
        t = t[1:]
    if not t or any(c not in "0123456789" for c in t):
        raise ValueError(s)
    v = -int(t) if neg else int(t)
    if v < lo or v > hi:
        raise ValueError(s)
    return v


def find_first_match_pos(cigar_str: str) -> int:    # utils.rs:12-42
    p = 0
    p_string = ""
    for c in cigar_str:
        if c not in OP_CHARS:
            p_string += c
        else:
            if c == "M":
                break
            if c in "SIX=":
                try:
                    p += _rust_parse_int(p_string, -(1 << 63), (1 << 63) - 1, True)
                except ValueError:
                    raise ReferencePanic("ffm parse")
            p_string = ""
    return p


def parse_cigar(cigar_str: str) -> dict:            # utils.rs:88-117
    m = {k: 0 for k in "DMIHSPX=N"}
    n_str = ""
    for x in cigar_str:
        if x in "0123456789":
            n_str += x
        else:
            if x not in OP_CHARS:
                # the reference opens a new bucket for any other char; outside the supported domain
                raise ReferencePanic("sa cigar char")
            try:
                n = _rust_parse_int(n_str, 0, U32, False)
            except ValueError:
                raise ReferencePanic("sa cigar count")
            m[x] = (m[x] + n) & U32
            n_str = ""
    return m


def parse_supplementary_alignment(s: str) -> Seg:   # utils.rs:119-139
    v = s.split(",")
    if len(v) < 6:
        raise ReferencePanic("sa fields")
    try:
        pos = _rust_parse_int(v[1], -(1 << 63), (1 << 63) - 1, True)
    except ValueError:
        raise ReferencePanic("sa pos")
    if v[2] == "+":
        strand = 1
    elif v[2] == "-":
        strand = -1
    else:
        raise ReferencePanic("sa strand")
    cm = parse_cigar(v[3])
    try:
        mapq = _rust_parse_int(v[4], 0, 255, False)
    except ValueError:
        raise ReferencePanic("sa mapq")
    try:
        _rust_parse_int(v[5], -(1 << 63), (1 << 63) - 1, True)
    except ValueError:
        raise ReferencePanic("sa nm")
    return _new_seg(v[0], pos - 1, cm, strand, mapq, v[3])


def overlap(a_start, a_end, b_start, b_end, max_over_pct) -> bool:      # utils.rs:158-194
    if a_end < b_start or a_start > b_end:
        return False
    min_len = min(a_end - a_start, b_end - b_start)

    def div(x, y):      # IEEE f64 division incl. x/0
        x = float(x); y = float(y)
        if y == 0.0:
            if x == 0.0:
                return float("nan")
            return float("inf") if x > 0 else float("-inf")
        return x / y

    if a_start < b_start:
        ov = div(a_end - b_start, min_len) if a_end < b_end else div(b_end - b_start, min_len)
    else:
        ov = div(b_end - a_start, min_len) if b_end < a_end else div(a_end - a_start, min_len)
    return ov > max_over_pct


@dataclass
class AEv:                      # AlignmentEvent, aligments_event.rs:11-24
    lchrom: str
    lstart: int
    lend: int
    lstrand: int
    rchrom: str
    rstart: int
    rend: int
    rstrand: int
    events_num: int
    is_del: bool


def _aev_new(chrom, left_consume, right_consume, event_len, pos, strand, is_del) -> AEv:    # aligments_event.rs:28-57
    c = _strip(chrom)
    pos2 = pos & U32
    return AEv(c, pos2, (pos2 + left_consume) & U32, strand, c,
               (pos2 + left_consume + event_len) & U32,
               (pos2 + left_consume + event_len + right_consume) & U32, strand, 1, is_del)


def _mk(a: AEv, b: AEv) -> AEv:
    return AEv(a.lchrom, a.lstart, a.lend, a.lstrand, b.rchrom, b.rstart, b.rend, b.rstrand, 1, True)


def _fmt_aev(x: AEv, P: Params, rec: Record, strand: int, tag: str) -> str:   # utils.rs:196-239
    base = f"{x.lchrom}\t{x.lstart}\t{x.lend}\t{x.lstrand}\t{x.rchrom}\t{x.rstart}\t{x.rend}\t{x.rstrand}\t{x.events_num}"
    if P.verbose:
        return base + f"\t{tag}\t{rec.qname}\tstrand:{strand}\tflag:{rec.flag}\n"
    return base + "\n"


def _fmt_split(a: Seg, b: Seg, P: Params, rec: Record, strand: int, n: int, tag: str) -> str:  # utils.rs:241-283
    base = f"{a.chrom}\t{a.start}\t{a.end}\t{a.strand}\t{b.chrom}\t{b.start}\t{b.end}\t{b.strand}\t{n - 1}"
    if P.verbose:
        return base + f"\t{tag}\t{rec.qname}\tstrand:{strand}\tflag:{rec.flag}\n"
    return base + "\n"


def merge_events(ev: List[AEv], merge_min: int) -> List[AEv]:     # main.rs:609-755
    """Literal merge; raises ReferencePanic where the Rust indexes out of bounds."""
    if len(ev) == 2:
        a, b = ev
        if abs(b.lend - a.rstart) < merge_min and a.is_del and b.is_del:
            return [_mk(a, b)]
        return [a, b]
    if len(ev) > 2:
        m1 = list(ev)
        m2: List[AEv] = []
        it = 5
        while True:
            it -= 1
            idx = 1
            while True:
                lim = (len(m1) - 2) & 0xFFFFFFFFFFFFFFFF          # usize wrap, main.rs:664
                if idx > lim:
                    break
                if idx + 1 >= len(m1):
                    raise ReferencePanic("merge index out of bounds")
                prv, tgt, nxt = m1[idx - 1], m1[idx], m1[idx + 1]
                mp = abs(prv.lend - tgt.rstart) < merge_min and tgt.is_del and prv.is_del
                mn = abs(tgt.lend - nxt.rstart) < merge_min and tgt.is_del and nxt.is_del
                if mp or mn:
                    if mp:
                        m2.append(_mk(prv, tgt))
                    if mn:
                        m2.append(_mk(tgt, nxt))
                else:
                    if idx == 1:
                        m2 += [prv, tgt, nxt]
                    else:
                        m2.append(nxt)
                idx += 1
            if len(m1) == len(m2):
                break
            elif it <= 0:
                break
            else:
                m1 = list(m2)
                m2 = []
        return list(m2)
    return list(ev)


def process_record(rec: Record, P: Params) -> List[str]:
    """One iteration of the loop at main.rs:158-770 -> the lines it writes, in order."""
    out: List[str] = []
    # filters main.rs:169-190
    if P.exclude_secondary and (rec.flag & 0x100):
        return out
    if P.exclude_unmapped and (rec.flag & 0x4):
        return out
    if rec.mapq < P.mapq:
        return out
    if rec.flag & P.exclude_flag:
        return out
    strand = -1 if rec.flag & 0x10 else 1
    if rec.contig is None:
        raise ReferencePanic("tid")
    contig_name, pos = rec.contig, rec.pos

    if rec.sa is not None:                                      # main.rs:206
        cigar_map = {k: 0 for k in "DMIHSPX=N"}
        first = ""
        for op, n in rec.cigar:                                 # main.rs:243-296
            if op > 8:
                raise ReferencePanic("cigar op")
            c = BAM_OPS[op]
            first += str(n) + c
            cigar_map[c] = (cigar_map[c] + n) & U32
        vec = [_new_seg(contig_name, pos, cigar_map, strand, rec.mapq, first)]
        if rec.sa_is_string:                                    # main.rs:308-320
            sa_list = rec.sa.split(";")
            if len(sa_list) > P.max_supp_alignm:
                return []                                       # `continue`: nothing at all for this record
            for piece in sa_list:
                if len(piece) > 0:
                    vec.append(parse_supplementary_alignment(piece))
        vec.sort(key=lambda s: find_first_match_pos(s.raw_cigar))   # stable, main.rs:322
        if len(vec) == 2:                                       # main.rs:340-451
            a, b = vec
            if a.cigar_map["S"] > P.ins_clip_min or a.cigar_map["H"] > P.ins_clip_min:
                if a.chrom == b.chrom:
                    if a.strand == b.strand:
                        if overlap(a.start, a.end, b.start, b.end, P.max_pct_overlap):
                            if b.cigar_map["S"] > P.ins_clip_min or b.cigar_map["H"] > P.ins_clip_min:
                                q = sorted([a.start, a.end, b.start, b.end])
                                for k in (1, 2):
                                    x = AEv(a.chrom, q[0] & U32, q[k] & U32, a.strand, b.chrom, q[k] & U32, q[k] & U32,
                                            b.strand, 1, False)
                                    out.append(_fmt_aev(x, P, rec, strand,
                                                        "excord-lr-alignment-event-large-ins-two-alignments"))
                else:
                    x = AEv(a.chrom, a.start & U32, a.end & U32, a.strand, a.chrom, a.end & U32, a.end & U32,
                            a.strand, 1, False)
                    out.append(_fmt_aev(x, P, rec, strand, "excord-lr-alignment-event-large-ins-one-alignments"))
        if len(vec) == 1:                                       # main.rs:459-486
            a = vec[0]
            if a.cigar_map["S"] > P.ins_clip_min or a.cigar_map["H"] > P.ins_clip_min:
                x = AEv(a.chrom, a.start & U32, a.end & U32, a.strand, a.chrom, a.end & U32, a.end & U32,
                        a.strand, 1, False)
                out.append(_fmt_aev(x, P, rec, strand, "excord-lr-alignment-event-large-ins"))
        for i in range(1, len(vec)):                            # main.rs:488-516
            a, b = vec[i - 1], vec[i]
            ka = (a.chrom.encode("latin-1"), a.start)           # alignment_pos_cmp, utils.rs:75-86
            kb = (b.chrom.encode("latin-1"), b.start)
            if ka > kb:
                a, b = b, a
            out.append(_fmt_split(a, b, P, rec, strand, len(vec), "excord-lr-split-read"))

    if not P.split_only:                                        # main.rs:523-768
        total = 0
        for op, n in rec.cigar:                                 # main.rs:528-545
            if op > 8:
                raise ReferencePanic("cigar op")
            if BAM_OPS[op] in "DMN=":
                total = (total + n) & U32
        left, right = 0, total
        ev: List[AEv] = []
        for op, n in rec.cigar:                                 # main.rs:549-600
            c = BAM_OPS[op]
            if c == "D":
                right = (right - n) & U32
                if n >= P.indel_min:
                    ev.append(_aev_new(contig_name, left, right, n, pos, strand, True))
                left = (left + n) & U32
            elif c == "I":
                if n >= P.indel_min:
                    ev.append(_aev_new(contig_name, left, n, 0, pos, strand, False))
            elif c in "MN=":
                left = (left + n) & U32
                right = (right - n) & U32
        try:
            merged = merge_events(ev, P.merge_min)
        except ReferencePanic as e:
            e.written = "".join(out)        # the SA arm's lines of this record are already in the BufWriter (main.rs:395-515)
            raise
        for x in merged:
            out.append(_fmt_aev(x, P, rec, strand, "excord-lr-alignment-event"))
    return out


def run(records: Sequence[Record], P: Params) -> str:
    """The whole output file for `records` in order; raises ReferencePanic(read=i)."""
    parts: List[str] = []
    for i, r in enumerate(records):
        try:
            parts.extend(process_record(r, P))
        except ReferencePanic as e:
            e.read = i
            e.partial = "".join(parts) + getattr(e, "written", "")
            raise
    return "".join(parts)
