"""Seeded adversarial record generator for parity tests (small batches, every code path)."""
import random

from excord_lr_b200.batch import SA_NONE, SA_OTHER, SA_STRING, ExlrParams, pack_records

REF_NAMES = ["chr1", "chr2", "1", "chr10", "chrX", "chrUn_KI270742v1", "HLA-A*01:01", "chr", "chrchr5", "20"]
SA_CHROMS = REF_NAMES + ["chr3", "7", "", "chrM"]


def rand_len(rng, big_ok=True):
    u = rng.random()
    if u < 0.70:
        return rng.randint(1, 40)
    if u < 0.90:
        return rng.randint(41, 120)
    if u < 0.985 or not big_ok:
        return rng.randint(121, 6000)
    return rng.choice([0, (1 << 28) - 1, (1 << 27), 65535, 1 << 20])


def rand_cigar_ops(rng, n_ops, style=None):
    style = style if style is not None else rng.choice([0, 0, 1, 2])
    if style == 0:
        codes = [0, 0, 0, 1, 2, 2, 4, 0, 1, 2]
    elif style == 1:
        codes = [7, 7, 8, 1, 2, 7, 8, 3, 4, 5]
    else:
        codes = list(range(9))
    return [(rand_len(rng) << 4) | rng.choice(codes) for _ in range(n_ops)]


def rand_cigar_text(rng):
    n = rng.randint(1, 8)
    ops = rand_cigar_ops(rng, n)
    s = "".join("%d%s" % (v >> 4, "MIDNSHP=X"[v & 15]) for v in ops)
    if rng.random() < 0.05:
        s += str(rng.randint(0, 99))           # trailing digits are ignored (utils.rs:104-114)
    if rng.random() < 0.05:
        s = "000" + s                          # leading zeros parse fine
    return s


def rand_sa(rng, max_pieces=6):
    n = rng.randint(0, max_pieces)
    parts = []
    for _ in range(n):
        pos = rng.choice([rng.randint(0, 250_000_000), rng.randint(0, 250_000_000), 0, 1, -5, 2 ** 40])
        pos_s = ("+" if rng.random() < 0.03 else "") + str(pos)
        fields = [rng.choice(SA_CHROMS), pos_s, rng.choice("+-"), rand_cigar_text(rng), str(rng.choice([0, 1, 60, 255])),
                  str(rng.randint(0, 500))]
        if rng.random() < 0.05:
            fields.append("extra")
        parts.append(",".join(fields))
    s = ";".join(parts)
    u = rng.random()
    if n and u < 0.8:
        s += ";"                               # the usual ';'-terminated form
    elif u > 0.97:
        s += ";;"                              # empty pieces count against -k but are skipped (main.rs:309-315)
    return s


def rand_record(rng, near=None):
    n_ops = rng.choice([0, 1, 2, 3, 5, 8, 13, 21, 34, 40]) if rng.random() < 0.9 else rng.randint(41, 700)
    r = dict(tid=rng.randrange(len(REF_NAMES)), pos=rng.choice([rng.randint(0, 2 ** 31 - 1), rng.randint(0, 10 ** 6), 0]),
             flag=rng.choice([0, 16, 0, 16, 2048, 2064, 256, 1024, 4, 512, 1, 3, 99, 147, rng.randrange(1 << 12)]),
             mapq=rng.choice([60, 60, 60, 1, 0, 255, 30]), cigar=rand_cigar_ops(rng, n_ops))
    u = rng.random()
    if u < 0.10:
        # large-INS candidates (main.rs:340-451): one SA piece, same/aliased chrom, nearby, big clips on both
        name = REF_NAMES[r["tid"]]
        alias = rng.choice([name, name, name[3:] if name.startswith("chr") else "chr" + name, rng.choice(SA_CHROMS)])
        clip = rng.choice([999, 1000, 1001, 30, 41, 5000])
        m = rng.randint(100, 3000)
        r["pos"] = rng.randint(10_000, 10 ** 6)
        r["cigar"] = [(rng.choice([0, 5, clip]) << 4) | 4, (m << 4) | 0, (clip << 4) | rng.choice([4, 5])]
        r["cigar"] = [v for v in r["cigar"] if v >> 4]
        st = "-" if r["flag"] & 16 else "+"
        if rng.random() < 0.15:
            st = "+" if st == "-" else "-"
        spos = r["pos"] + rng.randint(-m - 200, m + 200) + 1
        scig = "%d%s%dM%s" % (rng.choice([clip, 1001, 5000]), rng.choice("SH"), rng.randint(1, 3000),
                              rng.choice(["", "14S", "2000H"]))
        r["sa"] = "%s,%d,%s,%s,60,5;" % (alias, spos, st, scig)
    elif u < 0.35:
        r["sa"] = rand_sa(rng)
    elif u < 0.38:
        r["sa"] = ""
        r["sa_kind"] = SA_OTHER
    elif u < 0.40:
        r["sa"] = ""                           # an empty Z string: one (empty) piece
    return r


def clustered_dels(rng, n_events, gap_max, min_len=50):
    """CIGAR with n_events deletions separated by short matches (merge rules, main.rs:609-755)."""
    ops = [(rng.randint(50, 500) << 4) | 0]
    for i in range(n_events):
        op = 2 if rng.random() < 0.85 else 1
        ops.append((rng.randint(min_len, min_len + 60) << 4) | op)
        if i + 1 < n_events:
            g = rng.randint(0, gap_max)
            if g:
                ops.append((g << 4) | rng.choice([0, 0, 7, 8, 3]))
    ops.append((rng.randint(50, 500) << 4) | 0)
    return ops


def rand_batch(seed, n=200, merge_clusters=True, qnames=False):
    rng = random.Random(seed)
    recs = []
    for i in range(n):
        r = rand_record(rng)
        if merge_clusters and rng.random() < 0.15:
            r["cigar"] = clustered_dels(rng, rng.choice([2, 2, 2, 3, 4, 6]), rng.choice([0, 3, 4, 5, 6, 12]))
        if qnames:
            r["qname"] = "q%d/%d" % (seed, i)
        recs.append(r)
    return pack_records(recs, REF_NAMES)


def rand_params(seed):
    rng = random.Random(seed * 7919 + 13)
    return ExlrParams.make(mapq=rng.choice([0, 1, 1, 30, 61]), exclude_flag=rng.choice([1796, 1796, 0, 4, 2048, 3844]),
                           exclude_secondary=rng.random() < 0.2, exclude_unmapped=rng.random() < 0.2,
                           indel_min=rng.choice([50, 50, 30, 1, 100, 41]), merge_min=rng.choice([5, 5, 0, 1, 6, 100]),
                           ins_clip_min=rng.choice([1000, 1000, 0, 40, 5000]), split_only=rng.random() < 0.2,
                           max_pct_overlap=rng.choice([0.0, 0.8, 0.5, 1.0, -1.0]),
                           max_supp_alignm=rng.choice([4, 4, 8, 0, 1, 2, 3, 100]))
