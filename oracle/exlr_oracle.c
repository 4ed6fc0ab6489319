/*
 * exlr_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of excord-lr's per-record loop body
 * (reference src/main.rs:158-770 with src/utils.rs, src/aligments_event.rs and
 * src/split_read_event.rs).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product library
 * (libexlr_cuda.so) never does.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and it cannot
 * be built here (no cargo/rustc, rust-htslib unpinned and absent).  This file is pinned
 * instead against (a) the comment examples the reference does contain (utils.rs:50-57,
 * main.rs:357-365), (b) the hand-derived known answers of SURVEY.md Appendix B, and
 * (c) an independent Python restatement (oracle/oracle.py) on randomized records.
 *
 * It deliberately keeps the reference's structure (a text CIGAR + 9-bucket map per SA
 * record, a comparator that re-parses the text on every comparison, two passes over the
 * CIGAR for the indel arm, the literal 5-round merge loop) so that timing it says
 * something about the reference; it does not try to be a fast CPU implementation.
 *
 * Input is the same structure-of-arrays batch the C ABI (include/exlr.h) takes; output is
 * the same exlr_event records, so the GPU path is compared field by field and, through
 * exlr_oracle_format(), byte by byte.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/exlr.h"

/* bucket order of the reference's HashMap keys (main.rs:214-224, utils.rs:92-102) */
enum { B_D = 0, B_M, B_I, B_H, B_S, B_P, B_X, B_EQ, B_N, B_COUNT };
/* BAM op code -> char (rust-htslib Cigar enum order quoted in main.rs:226-234: M I D N S H P = X) */
static const char BAM_OPS[9] = { 'M', 'I', 'D', 'N', 'S', 'H', 'P', '=', 'X' };

static int bucket_of(char c)
{
    switch (c) {
    case 'D': return B_D;  case 'M': return B_M;  case 'I': return B_I;
    case 'H': return B_H;  case 'S': return B_S;  case 'P': return B_P;
    case 'X': return B_X;  case '=': return B_EQ; case 'N': return B_N;
    default:  return -1;
    }
}

/* SplitReadEvent (split_read_event.rs:3-11) */
typedef struct {
    uint32_t chrom_ref;     /* how the product names the chrom: tid, or sa offset | 1<<31 */
    const char* chrom;      /* "chr"-stripped name bytes */
    size_t chrom_len;
    int64_t start, end;
    uint32_t map[B_COUNT];
    int strand;
    uint8_t mapq;
    char* raw_cigar;        /* owned, NUL terminated */
} seg_t;

typedef struct exlr_oracle_out {
    exlr_event* events;
    uint64_t n_events, cap_events;
    uint32_t* line_off;     /* [n_reads+1] filled for [r_begin, r_end] relative to this run */
    int32_t status;
    uint32_t err_read;
    uint64_t n_kept, n_sa_reads, n_cap_dropped, n_ops;
    uint64_t n_err_lines;   /* lines the failing record had already written when it panicked (its SA-arm lines when the
                               indel arm panics: f.write at main.rs:395-515 precedes main.rs:523-742) */
} exlr_oracle_out;

static void push_event(exlr_oracle_out* o, const exlr_event* e)
{
    if (o->n_events == o->cap_events) {
        o->cap_events = o->cap_events ? o->cap_events * 2 : 1024;
        o->events = (exlr_event*)realloc(o->events, o->cap_events * sizeof(exlr_event));
        if (!o->events) { fprintf(stderr, "exlr_oracle: out of memory\n"); abort(); }
    }
    o->events[o->n_events++] = *e;
}

/* strip one leading "chr" (aligments_event.rs:38-42, split_read_event.rs:30-34) */
static void strip_chr(const char** s, size_t* n)
{
    if (*n >= 3 && (*s)[0] == 'c' && (*s)[1] == 'h' && (*s)[2] == 'r') { *s += 3; *n -= 3; }
}

/* SplitReadEvent::new (split_read_event.rs:14-45): end = start + D + M + '=' + X - 1, stored end + 1 */
static void seg_finish(seg_t* s)
{
    int64_t end = s->start + (int64_t)s->map[B_D] + (int64_t)s->map[B_M] + (int64_t)s->map[B_EQ]
                + (int64_t)s->map[B_X] + -1;
    s->end = end + 1;
}

/* Rust str::parse::<i64>: optional single sign, >= 1 ASCII digits, no overflow. */
static int parse_i64(const char* s, size_t n, int64_t* out)
{
    size_t i = 0; int neg = 0;
    if (n == 0) return 0;
    if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
    if (i == n) return 0;
    uint64_t lim = neg ? (uint64_t)1 << 63 : ((uint64_t)1 << 63) - 1, v = 0;
    for (; i < n; i++) {
        if (s[i] < '0' || s[i] > '9') return 0;
        unsigned d = (unsigned)(s[i] - '0');
        if (v > (lim - d) / 10) return 0;
        v = v * 10 + d;
    }
    *out = neg ? (int64_t)(0 - v) : (int64_t)v;
    return 1;
}

/* Rust str::parse::<u8>/<u32>: optional '+', >= 1 ASCII digits, value <= max. */
static int parse_unsigned(const char* s, size_t n, uint64_t max, uint64_t* out)
{
    size_t i = 0; uint64_t v = 0;
    if (n == 0) return 0;
    if (s[0] == '+') i = 1;
    if (i == n) return 0;
    for (; i < n; i++) {
        if (s[i] < '0' || s[i] > '9') return 0;
        v = v * 10 + (unsigned)(s[i] - '0');
        if (v > max) return 0;
    }
    *out = v;
    return 1;
}

/* parse_cigar (utils.rs:88-117).  Domain: bytes in [0-9MIDNSHP=X]; anything else is
 * reported as EXLR_ERR_SA_CIGAR (the reference would open an extra bucket). */
static int parse_cigar_text(const char* s, size_t n, uint32_t map[B_COUNT])
{
    size_t num_begin = 0;
    memset(map, 0, sizeof(uint32_t) * B_COUNT);
    for (size_t i = 0; i < n; i++) {
        char c = s[i];
        if (c >= '0' && c <= '9') continue;           /* n_str += x           (utils.rs:105-106) */
        int b = bucket_of(c);
        if (b < 0) return EXLR_ERR_SA_CIGAR;
        uint64_t v;
        if (i == num_begin) return EXLR_ERR_SA_CIGAR; /* "".parse::<u32>() -> unwrap panic (utils.rs:109) */
        if (!parse_unsigned(s + num_begin, i - num_begin, 0xffffffffull, &v)) return EXLR_ERR_SA_CIGAR;
        map[b] += (uint32_t)v;                        /* wrapping add in release (utils.rs:111) */
        num_begin = i + 1;
    }
    return EXLR_OK;                                   /* trailing digits are ignored */
}

/* find_first_match_pos (utils.rs:12-42) */
static int64_t find_first_match_pos(const char* cigar)
{
    int64_t p = 0;
    const char* num = cigar; size_t numlen = 0;
    for (const char* c = cigar; *c; c++) {
        if (bucket_of(*c) < 0) { if (numlen == 0) num = c; numlen++; }
        else {
            if (*c == 'M') break;
            if (*c == 'S' || *c == 'I' || *c == 'X' || *c == '=') {
                int64_t v = 0;
                parse_i64(num, numlen, &v);           /* validated by parse_cigar_text beforehand */
                p += v;
            }
            numlen = 0;
        }
    }
    return p;
}

/* splitter_order_cmp (utils.rs:58-73): recomputed on every comparison, as in the reference */
static int splitter_order_cmp(const seg_t* a, const seg_t* b)
{
    int64_t pa = find_first_match_pos(a->raw_cigar), pb = find_first_match_pos(b->raw_cigar);
    return pa < pb ? -1 : (pa > pb ? 1 : 0);
}

static int bytes_cmp(const char* a, size_t an, const char* b, size_t bn)
{
    size_t m = an < bn ? an : bn;
    int c = m ? memcmp(a, b, m) : 0;
    if (c) return c;
    return an < bn ? -1 : (an > bn ? 1 : 0);
}

/* alignment_pos_cmp (utils.rs:75-86) */
static int alignment_pos_cmp(const seg_t* a, const seg_t* b)
{
    int c = bytes_cmp(a->chrom, a->chrom_len, b->chrom, b->chrom_len);
    if (c) return c;
    return a->start < b->start ? -1 : (a->start > b->start ? 1 : 0);
}

/* overlap (utils.rs:158-194) */
static int overlap(int64_t a_start, int64_t a_end, int64_t b_start, int64_t b_end, double max_over_pct)
{
    if (a_end < b_start || a_start > b_end) return 0;
    int64_t la = a_end - a_start, lb = b_end - b_start;
    int64_t min_len = la < lb ? la : lb;
    double ov;
    if (a_start < b_start) {
        if (a_end < b_end) ov = (double)(a_end - b_start) / (double)min_len;
        else               ov = (double)(b_end - b_start) / (double)min_len;
    } else {
        if (b_end < a_end) ov = (double)(b_end - a_start) / (double)min_len;
        else               ov = (double)(a_end - a_start) / (double)min_len;
    }
    return ov > max_over_pct;
}

/* parse_supplementary_alignment (utils.rs:119-139).  Error priority (all are panics in the
 * reference; the order only fixes which diagnostic code is reported): FIELDS, POS, STRAND,
 * CIGAR, MAPQ, NM. */
static int parse_supplementary_alignment(const char* s, size_t n, uint32_t sa_byte_off, seg_t* out)
{
    const char* f[6]; size_t fl[6]; int nf = 0;
    size_t b = 0;
    for (size_t i = 0; i <= n; i++) {
        if (i == n || s[i] == ',') {
            if (nf < 6) { f[nf] = s + b; fl[nf] = i - b; }
            nf++; b = i + 1;
        }
    }
    if (nf < 6) return EXLR_ERR_SA_FIELDS;
    int64_t pos;
    if (!parse_i64(f[1], fl[1], &pos)) return EXLR_ERR_SA_POS;
    int strand;
    if (fl[2] == 1 && f[2][0] == '+') strand = 1;
    else if (fl[2] == 1 && f[2][0] == '-') strand = -1;
    else return EXLR_ERR_SA_STRAND;
    int rc = parse_cigar_text(f[3], fl[3], out->map);
    if (rc) return rc;
    uint64_t mq, dummy_ok; int64_t nm;
    if (!parse_unsigned(f[4], fl[4], 255, &mq)) return EXLR_ERR_SA_MAPQ;
    if (!parse_i64(f[5], fl[5], &nm)) return EXLR_ERR_SA_NM;
    (void)dummy_ok; (void)nm;
    const char* cs = f[0]; size_t cl = fl[0];
    strip_chr(&cs, &cl);
    out->chrom = cs; out->chrom_len = cl;
    out->chrom_ref = 0x80000000u | (sa_byte_off + (uint32_t)(cs - s));
    out->start = (int64_t)((uint64_t)pos - 1u);       /* pos - 1 (utils.rs:133) */
    out->strand = strand;
    out->mapq = (uint8_t)mq;
    out->raw_cigar = (char*)malloc(fl[3] + 1);
    memcpy(out->raw_cigar, f[3], fl[3]); out->raw_cigar[fl[3]] = 0;
    seg_finish(out);
    return EXLR_OK;
}

/* AlignmentEvent (aligments_event.rs:11-24), chrom kept as a tid */
typedef struct { uint32_t lstart, lend, rstart, rend; int is_del; } aev_t;

/* AlignmentEvent::new (aligments_event.rs:28-57): u32 wrapping arithmetic (release build) */
static aev_t aev_new(uint32_t left_consume, uint32_t right_consume, uint32_t event_len, int64_t pos, int is_del)
{
    aev_t e;
    uint32_t pos2 = (uint32_t)pos;                    /* *pos as u32 (aligments_event.rs:43) */
    e.lstart = pos2;
    e.lend = pos2 + left_consume;
    e.rstart = pos2 + left_consume + event_len;
    e.rend = pos2 + left_consume + event_len + right_consume;
    e.is_del = is_del;
    return e;
}

static uint32_t abs_diff_u32(uint32_t a, uint32_t b) { return a > b ? a - b : b - a; }

static aev_t aev_merge(aev_t a, aev_t b)
{
    aev_t c; c.lstart = a.lstart; c.lend = a.lend; c.rstart = b.rstart; c.rend = b.rend; c.is_del = 1;
    return c;
}

typedef struct { aev_t* v; size_t n, cap; } aev_vec;
static void aev_push(aev_vec* a, aev_t e)
{
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 8; a->v = (aev_t*)realloc(a->v, a->cap * sizeof(aev_t)); }
    a->v[a->n++] = e;
}

/* The literal >2 merge loop (main.rs:636-742).  Returns 0 ok, 1 if the reference would panic
 * (index out of bounds after merge1.len()-2 wraps, main.rs:664-671).  *changed is set when
 * any predicate fired (the result is then not the identity). */
static int merge_loop(const aev_vec* ev, uint32_t merge_min, aev_vec* result, int* changed)
{
    aev_vec m1 = {0}, m2 = {0};
    for (size_t i = 0; i < ev->n; i++) aev_push(&m1, ev->v[i]);
    uint32_t iter_times = 5;
    int panic = 0;
    *changed = 0;
    for (;;) {
        iter_times -= 1;
        size_t idx = 1;
        for (;;) {
            size_t lim = m1.n - 2;                     /* usize wrap when m1.n < 2 (main.rs:664) */
            if (idx > lim) break;
            size_t pidx = idx - 1, nidx = idx + 1;
            if (nidx >= m1.n) { panic = 1; break; }    /* merge1[nidx] out of bounds -> panic */
            aev_t prv = m1.v[pidx], target = m1.v[idx], nxt = m1.v[nidx];
            int mp = abs_diff_u32(prv.lend, target.rstart) < merge_min && (target.is_del && prv.is_del);
            int mn = abs_diff_u32(target.lend, nxt.rstart) < merge_min && (target.is_del && nxt.is_del);
            if (mp || mn) {
                *changed = 1;
                if (mp) aev_push(&m2, aev_merge(prv, target));
                if (mn) aev_push(&m2, aev_merge(target, nxt));
                idx += 1;
            } else {
                if (idx == 1) { aev_push(&m2, prv); aev_push(&m2, target); aev_push(&m2, nxt); }
                else aev_push(&m2, nxt);
                idx += 1;
            }
        }
        if (panic) break;
        if (m1.n == m2.n) break;
        else if (iter_times <= 0) break;
        else { m1.n = 0; for (size_t i = 0; i < m2.n; i++) aev_push(&m1, m2.v[i]); m2.n = 0; }
    }
    result->n = 0;
    for (size_t i = 0; i < m2.n; i++) aev_push(result, m2.v[i]);
    free(m1.v); free(m2.v);
    return panic;
}

static void emit_aev(exlr_oracle_out* o, uint32_t read, uint32_t lchrom, uint32_t rchrom,
                     uint32_t lstart, uint32_t lend, uint32_t rstart, uint32_t rend,
                     int lstrand, int rstrand, unsigned kind)
{
    exlr_event e;
    e.lstart = lstart; e.lend = lend; e.rstart = rstart; e.rend = rend;   /* zero-extended u32 */
    e.read_idx = read; e.lchrom = lchrom; e.rchrom = rchrom;
    e.meta = EXLR_EV_META(1, kind, lstrand < 0, rstrand < 0);
    push_event(o, &e);
}

/*
 * The >2 merge loop is literal (main.rs:636-742): non-identity results are emitted as the reference prints
 * them; only the panic case (index out of bounds) reports EXLR_ERR_MERGE_DOMAIN.  `merge_mode` is kept in the
 * signature for callers of round 1 and ignored.
 */
int exlr_oracle_run(const exlr_params* P, const char* const* ref_names, int n_ref,
                    uint64_t n_reads, const uint32_t* cigar, const uint64_t* cigar_off,
                    const int32_t* pos_a, const int32_t* tid_a, const uint16_t* flag_a,
                    const uint8_t* mapq_a, const uint8_t* sa_kind, const uint32_t* sa_off,
                    const uint8_t* sa_bytes, int merge_mode, uint64_t r_begin, uint64_t r_end,
                    exlr_oracle_out* o)
{
    memset(o, 0, sizeof(*o));
    if (r_end > n_reads) r_end = n_reads;
    if (r_begin > r_end) r_begin = r_end;
    o->line_off = (uint32_t*)calloc((size_t)(r_end - r_begin) + 1, sizeof(uint32_t));
    o->status = EXLR_OK; o->err_read = 0xffffffffu;

    seg_t* segs = NULL; size_t segs_cap = 0;
    aev_vec ev = {0}, merged = {0};
    char* first_cigar_str = NULL; size_t fcs_cap = 0;

#define FAIL(code) do { o->status = (code); o->err_read = (uint32_t)r; goto done; } while (0)

    for (uint64_t r = r_begin; r < r_end; r++) {
        o->line_off[r - r_begin] = (uint32_t)o->n_events;
        uint16_t flags = flag_a[r];
        const uint32_t* cg = cigar + cigar_off[r];
        size_t n_cigar = (size_t)(cigar_off[r + 1] - cigar_off[r]);
        o->n_ops += n_cigar;
        /* filters, main.rs:169-190 */
        if (P->exclude_secondary && (flags & 0x100)) continue;
        if (P->exclude_unmapped && (flags & 0x4)) continue;
        if (mapq_a[r] < P->mapq) continue;
        if ((flags & P->exclude_flag) != 0) continue;
        o->n_kept++;
        /* main.rs:196-203 */
        int strand = (flags & 0x10) ? -1 : 1;
        int32_t tid = tid_a[r];
        if (tid < 0 || tid >= n_ref) FAIL(EXLR_ERR_TID);        /* record.contig() panics */
        int64_t pos = pos_a[r];
        ev.n = 0;
        size_t nseg = 0;

        if (sa_kind[r] != EXLR_SA_NONE) {                         /* main.rs:206 Ok(_sa) */
            o->n_sa_reads++;
            uint32_t cigar_map[B_COUNT]; memset(cigar_map, 0, sizeof cigar_map);
            size_t need = n_cigar * 11 + 1, len = 0;
            if (need > fcs_cap) { fcs_cap = need * 2; first_cigar_str = (char*)realloc(first_cigar_str, fcs_cap); }
            for (size_t i = 0; i < n_cigar; i++) {                /* main.rs:243-296 */
                uint32_t opc = cg[i] & 0xf, n = cg[i] >> 4;
                if (opc > 8) FAIL(EXLR_ERR_CIGAR_OP);
                len += (size_t)sprintf(first_cigar_str + len, "%u", n);
                first_cigar_str[len++] = BAM_OPS[opc];
                cigar_map[bucket_of(BAM_OPS[opc])] += n;
            }
            first_cigar_str[len] = 0;
            if (segs_cap < 1) { segs_cap = 16; segs = (seg_t*)realloc(segs, segs_cap * sizeof(seg_t)); }
            seg_t* s0 = &segs[0];                                 /* main.rs:299-306 */
            const char* cn = ref_names[tid]; size_t cl = strlen(cn);
            strip_chr(&cn, &cl);
            s0->chrom = cn; s0->chrom_len = cl; s0->chrom_ref = (uint32_t)tid;
            s0->start = pos; memcpy(s0->map, cigar_map, sizeof cigar_map);
            s0->strand = strand; s0->mapq = mapq_a[r];
            s0->raw_cigar = (char*)malloc(len + 1); memcpy(s0->raw_cigar, first_cigar_str, len + 1);
            seg_finish(s0);
            nseg = 1;
            int skip_record = 0, sa_err = 0;
            if (sa_kind[r] == EXLR_SA_STRING) {                   /* main.rs:308-320 */
                const char* sa = (const char*)sa_bytes + sa_off[r];
                size_t sal = sa_off[r + 1] - sa_off[r];
                uint64_t pieces = 1;
                for (size_t i = 0; i < sal; i++) if (sa[i] == ';') pieces++;
                if (pieces > P->max_supp_alignm) skip_record = 1; /* main.rs:311-313 `continue` */
                else {
                    size_t b = 0;
                    for (size_t i = 0; i <= sal && !sa_err; i++) {
                        if (i == sal || sa[i] == ';') {
                            if (i > b) {                          /* filter(|x| x.len() > 0) */
                                if (nseg == segs_cap) { segs_cap *= 2; segs = (seg_t*)realloc(segs, segs_cap * sizeof(seg_t)); }
                                sa_err = parse_supplementary_alignment(sa + b, i - b, sa_off[r] + (uint32_t)b, &segs[nseg]);
                                if (!sa_err) nseg++;
                            }
                            b = i + 1;
                        }
                    }
                }
            }
            if (skip_record || sa_err) {
                for (size_t i = 0; i < nseg; i++) free(segs[i].raw_cigar);
                if (sa_err) FAIL(sa_err);
                o->n_cap_dropped++;
                continue;
            }
            /* alignment_vec.sort_by(splitter_order_cmp): stable (main.rs:322) */
            for (size_t i = 1; i < nseg; i++) {
                seg_t x = segs[i]; size_t j = i;
                while (j > 0 && splitter_order_cmp(&segs[j - 1], &x) > 0) { segs[j] = segs[j - 1]; j--; }
                segs[j] = x;
            }
            if (nseg - 1 >= (1u << 24)) {
                for (size_t i = 0; i < nseg; i++) free(segs[i].raw_cigar);
                FAIL(EXLR_ERR_SPLIT_COUNT);
            }
            if (nseg == 2) {                                      /* main.rs:340-451 */
                const seg_t* a = &segs[0]; const seg_t* b = &segs[1];
                if (a->map[B_S] > P->ins_clip_min || a->map[B_H] > P->ins_clip_min) {
                    if (bytes_cmp(a->chrom, a->chrom_len, b->chrom, b->chrom_len) == 0) {
                        if (a->strand == b->strand) {
                            if (overlap(a->start, a->end, b->start, b->end, P->max_pct_overlap)) {
                                if (b->map[B_S] > P->ins_clip_min || b->map[B_H] > P->ins_clip_min) {
                                    int64_t q[4] = { a->start, a->end, b->start, b->end };
                                    for (int i = 1; i < 4; i++) {   /* pos_list.sort() */
                                        int64_t x = q[i]; int j = i;
                                        while (j > 0 && q[j - 1] > x) { q[j] = q[j - 1]; j--; }
                                        q[j] = x;
                                    }
                                    emit_aev(o, (uint32_t)r, a->chrom_ref, b->chrom_ref, (uint32_t)q[0], (uint32_t)q[1],
                                             (uint32_t)q[1], (uint32_t)q[1], a->strand, b->strand, EXLR_KIND_INS_TWO_ALN);
                                    emit_aev(o, (uint32_t)r, a->chrom_ref, b->chrom_ref, (uint32_t)q[0], (uint32_t)q[2],
                                             (uint32_t)q[2], (uint32_t)q[2], a->strand, b->strand, EXLR_KIND_INS_TWO_ALN);
                                }
                            }
                        }
                    } else {
                        emit_aev(o, (uint32_t)r, a->chrom_ref, a->chrom_ref, (uint32_t)a->start, (uint32_t)a->end,
                                 (uint32_t)a->end, (uint32_t)a->end, a->strand, a->strand, EXLR_KIND_INS_ONE_ALN);
                    }
                }
            }
            if (nseg == 1) {                                      /* main.rs:459-486 */
                const seg_t* a = &segs[0];
                if (a->map[B_S] > P->ins_clip_min || a->map[B_H] > P->ins_clip_min)
                    emit_aev(o, (uint32_t)r, a->chrom_ref, a->chrom_ref, (uint32_t)a->start, (uint32_t)a->end,
                             (uint32_t)a->end, (uint32_t)a->end, a->strand, a->strand, EXLR_KIND_INS_ONE_SEG);
            }
            for (size_t i = 1; i < nseg; i++) {                   /* main.rs:488-516 */
                const seg_t* a = &segs[i - 1]; const seg_t* b = &segs[i];
                if (alignment_pos_cmp(a, b) > 0) { const seg_t* t = a; a = b; b = t; }
                exlr_event e;
                e.lstart = a->start; e.lend = a->end; e.rstart = b->start; e.rend = b->end;
                e.read_idx = (uint32_t)r; e.lchrom = a->chrom_ref; e.rchrom = b->chrom_ref;
                e.meta = EXLR_EV_META(nseg - 1, EXLR_KIND_SPLIT, a->strand < 0, b->strand < 0);
                push_event(o, &e);
            }
            for (size_t i = 0; i < nseg; i++) free(segs[i].raw_cigar);
        }

        if (!P->split_only) {                                     /* main.rs:523-768 */
            uint32_t total_consume = 0;
            for (size_t i = 0; i < n_cigar; i++) {                /* main.rs:528-545 */
                uint32_t opc = cg[i] & 0xf, n = cg[i] >> 4;
                if (opc > 8) FAIL(EXLR_ERR_CIGAR_OP);
                char c = BAM_OPS[opc];
                if (c == 'D' || c == 'M' || c == 'N' || c == '=') total_consume += n;
            }
            uint32_t left_consume = 0, right_consume = total_consume;
            for (size_t i = 0; i < n_cigar; i++) {                /* main.rs:549-600 */
                uint32_t opc = cg[i] & 0xf, n = cg[i] >> 4;
                char c = BAM_OPS[opc];
                if (c == 'D') {
                    right_consume -= n;
                    if (n >= P->indel_min) aev_push(&ev, aev_new(left_consume, right_consume, n, pos, 1));
                    left_consume += n;
                } else if (c == 'I') {
                    if (n >= P->indel_min) aev_push(&ev, aev_new(left_consume, n, 0u, pos, 0));
                } else if (c == 'M' || c == 'N' || c == '=') {
                    left_consume += n; right_consume -= n;
                }
            }
            merged.n = 0;                                         /* main.rs:609-755 */
            if (ev.n == 2) {
                aev_t a = ev.v[0], b = ev.v[1];
                if (abs_diff_u32(b.lend, a.rstart) < P->merge_min && (a.is_del && b.is_del)) aev_push(&merged, aev_merge(a, b));
                else { aev_push(&merged, a); aev_push(&merged, b); }
            } else if (ev.n > 2) {
                int changed = 0;
                int panic = merge_loop(&ev, P->merge_min, &merged, &changed);
                (void)changed; (void)merge_mode;
                if (panic) FAIL(EXLR_ERR_MERGE_DOMAIN);
            } else if (ev.n == 1) {
                aev_push(&merged, ev.v[0]);
            }
            for (size_t i = 0; i < merged.n; i++)                 /* main.rs:757-767 */
                emit_aev(o, (uint32_t)r, (uint32_t)tid, (uint32_t)tid, merged.v[i].lstart, merged.v[i].lend,
                         merged.v[i].rstart, merged.v[i].rend, strand, strand, EXLR_KIND_INDEL);
        }
    }
done:
    {
        uint64_t stop = (o->status == EXLR_OK) ? r_end : (uint64_t)o->err_read;
        /* The lines a failing record wrote before it panicked stay in the BufWriter and are flushed on unwind.  Only a
         * panic in the indel arm comes after any write of the same record (the SA arm's lines, main.rs:395-515); every
         * other panic (contig(), the CIGAR decode at :243, the SA parse at :315-320) precedes the record's first write. */
        if (o->status != EXLR_OK) {
            if (o->status == EXLR_ERR_MERGE_DOMAIN) o->n_err_lines = o->n_events - o->line_off[stop - r_begin];
            else o->n_events = o->line_off[stop - r_begin];
        }
        for (uint64_t r = stop + (o->status != EXLR_OK ? 1 : 0); r <= r_end; r++) o->line_off[r - r_begin] = (uint32_t)o->n_events;
    }
    free(segs); free(ev.v); free(merged.v); free(first_cigar_str);
    return o->status;
#undef FAIL
}

void exlr_oracle_free(exlr_oracle_out* o)
{
    free(o->events); free(o->line_off);
    memset(o, 0, sizeof(*o));
}

static const char* kind_tag(unsigned kind)
{
    switch (kind) {
    case EXLR_KIND_INDEL:       return "excord-lr-alignment-event";
    case EXLR_KIND_INS_ONE_SEG: return "excord-lr-alignment-event-large-ins";
    case EXLR_KIND_INS_ONE_ALN: return "excord-lr-alignment-event-large-ins-one-alignments";
    case EXLR_KIND_INS_TWO_ALN: return "excord-lr-alignment-event-large-ins-two-alignments";
    default:                    return "excord-lr-split-read";
    }
}

/* get_alignment_event_record / get_alignment_split_record (utils.rs:196-283).
 * Returns bytes needed; writes at most out_cap. */
int64_t exlr_oracle_format(const exlr_event* ev, uint64_t n_ev, const char* const* ref_names,
                           const uint8_t* sa_bytes, const uint16_t* flag_a, int verbose,
                           const char* qnames, const uint32_t* qname_off, char* out, uint64_t out_cap)
{
    uint64_t w = 0;
    char* line = NULL; size_t line_cap = 0;
    for (uint64_t i = 0; i < n_ev; i++) {
        const exlr_event* e = &ev[i];
        const char* cs[2]; size_t cl[2];
        uint32_t refs[2] = { e->lchrom, e->rchrom };
        for (int s = 0; s < 2; s++) {
            if (EXLR_CHROM_IS_SA(refs[s])) {
                const char* p = (const char*)sa_bytes + EXLR_CHROM_SA_OFF(refs[s]);
                size_t n = 0; while (p[n] != ',') n++;
                cs[s] = p; cl[s] = n;
            } else {
                const char* p = ref_names[refs[s]]; size_t n = strlen(p);
                strip_chr(&p, &n);
                cs[s] = p; cl[s] = n;
            }
        }
        unsigned kind = EXLR_EV_KIND(e->meta);
        uint32_t r = e->read_idx;
        size_t qn = verbose ? (size_t)(qname_off[r + 1] - qname_off[r]) : 0;
        size_t need = cl[0] + cl[1] + qn + 320;
        if (need > line_cap) { line_cap = need * 2; line = (char*)realloc(line, line_cap); }
        int n = snprintf(line, line_cap, "%.*s\t%lld\t%lld\t%d\t%.*s\t%lld\t%lld\t%d\t%u",
                         (int)cl[0], cs[0], (long long)e->lstart, (long long)e->lend, EXLR_EV_LSTRAND(e->meta),
                         (int)cl[1], cs[1], (long long)e->rstart, (long long)e->rend, EXLR_EV_RSTRAND(e->meta),
                         EXLR_EV_NUM(e->meta));
        if (verbose) {
            int strand = (flag_a[r] & 0x10) ? -1 : 1;
            n += snprintf(line + n, line_cap - (size_t)n, "\t%s\t%.*s\tstrand:%d\tflag:%u", kind_tag(kind),
                          (int)qn, qnames + qname_off[r], strand, (unsigned)flag_a[r]);
        }
        line[n++] = '\n';
        if (w + (uint64_t)n <= out_cap) memcpy(out + w, line, (size_t)n);
        w += (uint64_t)n;
    }
    free(line);
    return (int64_t)w;
}
