// exlr_sa.cu — the SA arm (reference src/main.rs:206-519, src/utils.rs, src/split_read_event.rs):
//
//   kernel 3a k3a_sa_cigar   per-op-type sums + first-match offset of SA records' own CIGARs
//                            (main.rs:214-306, utils.rs:12-42), 2/4/8-lane group per record
//   kernel 3b k3b_sa_events  SA parse (utils.rs:88-139), -k cap (main.rs:311), stable segment sort
//                            (main.rs:322), large-INS rules (:340-486), split pairs (:488-516)
#include "exlr_common.cuh"
#include "exlr_sa_parse.cuh"

namespace exlr {

// ======================================================================================
// kernel 3a: SA records' own CIGAR -> clip sums, reference span, first-match offset
// ======================================================================================
// One record's CIGAR walked by a group of G lanes (G = 8: a quarter warp per short record; G = 32: the whole warp for a
// long one).  Returns the reduced sums in every lane of the group.
struct K3aAcc { uint32_t S, H, D, M, E, X; unsigned long long ffm; };

template <int G>
__device__ __forceinline__ K3aAcc k3a_walk(const DevBatch& B, uint32_t r, unsigned long long o0, unsigned long long o1, uint32_t gmask, uint32_t gshift)
{
    const uint32_t sub = threadIdx.x & (G - 1);
    K3aAcc a{0, 0, 0, 0, 0, 0, 0};
    bool seenM = false;
    constexpr int U = G == 32 ? 8 : (G == 8 ? 4 : 8);                          // independent loads in flight per lane
    for (unsigned long long b = o0; b < o1; b += (unsigned long long)G * U) {
        uint32_t vv[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const unsigned long long at = b + (unsigned long long)u * G + sub;
            vv[u] = at < o1 ? __ldg(B.cigar + at) : 0xfu;                     // 0xf: not an op
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool valid = b + (unsigned long long)u * G + sub < o1;
            const uint32_t v = vv[u];
            const uint32_t op = v & 15u, len = v >> 4;
            if (valid && op > 8u) report(B.ctrl, r, RANK_CIGAR_OP);
            a.M += op == 0u ? len : 0u; a.D += op == 2u ? len : 0u; a.S += op == 4u ? len : 0u;
            a.H += op == 5u ? len : 0u; a.E += op == 7u ? len : 0u; a.X += op == 8u ? len : 0u;
            if (!seenM) {                                                      // utils.rs:28-29 stops at the first M
                const uint32_t mb = (__ballot_sync(gmask, valid && op == 0u) >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
                const bool before = (mb & ((2u << sub) - 1u)) == 0u;          // no M at or before my lane
                if (valid && before && (op == 4u || op == 1u || op == 8u || op == 7u)) a.ffm += len;   // S I X =  (utils.rs:33)
                if (mb) seenM = true;
            }
        }
    }
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        a.S += __shfl_xor_sync(gmask, a.S, d); a.H += __shfl_xor_sync(gmask, a.H, d); a.D += __shfl_xor_sync(gmask, a.D, d);
        a.M += __shfl_xor_sync(gmask, a.M, d); a.E += __shfl_xor_sync(gmask, a.E, d); a.X += __shfl_xor_sync(gmask, a.X, d);
        a.ffm += __shfl_xor_sync(gmask, a.ffm, d);
    }
    return a;
}

__device__ __forceinline__ void k3a_store(const DevBatch& B, uint32_t j, const K3aAcc& a)
{
    SaSum o;
    o.S = a.S; o.H = a.H;
    o.refspan = (int64_t)a.D + (int64_t)a.M + (int64_t)a.E + (int64_t)a.X;
    o.ffm = (int64_t)a.ffm; o.pad[0] = o.pad[1] = 0;
    B.sa_sum[j] = o;
}

static constexpr uint32_t K3A_LONG = 96;       // CIGARs longer than this are walked by the whole warp

// Each warp takes 32/G consecutive SA-list entries: short CIGARs are walked by the G-lane groups side by side, long ones
// (ONT: 10^3..10^5 ops) by all 32 lanes one after the other, so the loads stay 128-byte coalesced.  G is picked by the host
// from the batch's mean CIGAR length (2 lanes for the few-op records of split-heavy batches, 8 for HiFi-like CIGARs).
template <int G>
__global__ void __launch_bounds__(256) k3a_sa_cigar(DevBatch B, DevParams P)
{
    constexpr uint32_t PER_WARP = 32 / G;
    griddep_wait();                                    // kernel 0's list and count
    griddep_launch();
    CtaTrace tr(B, 3);
    const uint32_t n_sa = B.ctrl->n_sa;
    tr.mid();
    const uint32_t lane = threadIdx.x & 31, sub = lane & (G - 1), grp = lane / G;
    const uint32_t gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (grp * G);
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t j0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * PER_WARP; j0 < n_sa; j0 += warps * PER_WARP) {
        const uint32_t j = j0 + grp;
        const bool have = j < n_sa;
        uint32_t r = 0; unsigned long long o0 = 0, o1 = 0;
        if (have) { r = B.sa_list[j]; o0 = B.cigar_off[r]; o1 = B.cigar_off[r + 1]; }
        const bool is_long = have && (o1 - o0) > K3A_LONG;
        if (have && !is_long) {
            const K3aAcc a = k3a_walk<G>(B, r, o0, o1, gmask, grp * G);
            if (sub == 0) k3a_store(B, j, a);
        }
        uint32_t longs = __ballot_sync(0xffffffffu, is_long && sub == 0);     // one bit per group (its lane 0)
        while (longs) {
            const int src = __ffs(longs) - 1; longs &= longs - 1;
            const uint32_t rr = __shfl_sync(0xffffffffu, r, src);
            const unsigned long long a0 = __shfl_sync(0xffffffffu, o0, src), a1 = __shfl_sync(0xffffffffu, o1, src);
            const K3aAcc a = k3a_walk<32>(B, rr, a0, a1, 0xffffffffu, 0);
            if (lane == 0) k3a_store(B, j0 + (uint32_t)src / G, a);
        }
    }
    tr.end();
}

// ======================================================================================
// kernel 3b: SA parse, cap, sort, large-INS rules, split pairs
//
// The SA list is ordered and only SA records own SA bytes, so the SA strings of 128 consecutive
// list entries are one contiguous byte range: a CTA stages it into shared memory with coalesced
// 128-bit loads and each thread then parses its own record's string out of shared memory
// (falls back to reading global memory when the range does not fit).
// ======================================================================================
static constexpr int K3B_THREADS = 192;                 // the usual tile (64 records, ~140 SA pieces) parses in one round
static constexpr int K3B_TILE = 64;                     // SA records per tile (pieces are then spread over all threads)
static constexpr int K3B_CTAS = 7;                      // CTAs per SM: 29 KB of shared memory and 48 registers x 192 threads each
static constexpr uint32_t K3B_STAGE_BYTES = 12 * 1024;
static constexpr uint32_t K3B_MAXP = 320;              // segments (records + SA pieces) per tile in the staged layout

struct SmemBytes {            // byte i of sa_bytes, served from the staged copy (bias is a multiple of 16)
    static constexpr bool kWords = true;
    const uint8_t* p; uint32_t bias;
    __device__ __forceinline__ uint32_t operator[](uint32_t i) const { return p[i - bias]; }
    // the aligned 32-bit word holding byte i (little endian: byte i is bits 8*(i&3)..)
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return *reinterpret_cast<const uint32_t*>(p + ((i - bias) & ~3u)); }
};
struct GlobalBytes {
    static constexpr bool kWords = false;
    const uint8_t* p;
    __device__ __forceinline__ uint32_t operator[](uint32_t i) const { return __ldg(p + i); }
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return 0u; }
};

// SWAR: 0x80 in every byte of x that equals c (exact: no carries cross byte lanes)
__device__ __forceinline__ uint32_t swar_eq(uint32_t x, uint32_t c4)
{
    const uint32_t y = x ^ c4;
    return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y | 0x7f7f7f7fu);
}
// the four "byte == ';'" flags of a word as a nibble (bit k: byte k)
__device__ __forceinline__ uint32_t semi_bits(uint32_t w) { return (((swar_eq(w, 0x3b3b3b3bu) >> 7) * 0x00204081u) >> 21) & 0xfu; }
// keep only the flag bits of the bytes whose absolute index is in [lo, hi); w0 = absolute index of the word's byte 0
__device__ __forceinline__ uint32_t swar_clip(uint32_t z, uint32_t w0, uint32_t lo, uint32_t hi)
{
    if (w0 >= lo && w0 + 4u <= hi) return z;             // interior word: the common case costs one predicate
    if (w0 < lo) z &= 0xffffffffu << (8u * (lo - w0));
    if (w0 + 4u > hi) z &= 0xffffffffu >> (8u * (w0 + 4u - hi));
    return z;
}

template <class Bytes>
__device__ __forceinline__ bool dev_parse_i64(const Bytes& s, uint32_t b, uint32_t e, int64_t* out)
{
    if (b == e) return false;
    bool neg = false;
    const uint32_t c0 = s[b];
    if (c0 == '+' || c0 == '-') { neg = c0 == '-'; b++; }
    if (b == e) return false;
    unsigned long long v = 0;
    if (e - b <= 18u) {                                  // < 10^18: cannot overflow an i64, no per-digit range check
        for (; b < e; b++) {
            const uint32_t d = s[b] - '0';
            if (d > 9u) return false;
            v = v * 10ull + d;
        }
    } else {
        const unsigned long long lim = neg ? (1ull << 63) : (1ull << 63) - 1;
        for (; b < e; b++) {
            const uint32_t d = s[b] - '0';
            if (d > 9u) return false;
            if (v > (lim - d) / 10ull) return false;
            v = v * 10ull + d;
        }
    }
    *out = neg ? (int64_t)(0ull - v) : (int64_t)v;
    return true;
}

template <class Bytes>
__device__ __forceinline__ bool dev_parse_u8(const Bytes& s, uint32_t b, uint32_t e)
{
    if (b == e) return false;
    if (s[b] == '+') b++;
    if (b == e) return false;
    uint32_t v = 0;
    for (; b < e; b++) {
        const uint32_t d = s[b] - '0';
        if (d > 9u) return false;
        v = v * 10u + d;
        if (v > 255u) return false;
    }
    return true;
}

// parse_supplementary_alignment + parse_cigar + find_first_match_pos (utils.rs:12-42, 88-139)
template <class Bytes>
__device__ uint32_t dev_parse_piece(const Bytes& s, uint32_t b, uint32_t e, const DevParams& P, Seg* out)
{
    uint32_t fb[6], fe[6], nf = 0, st = b;
    if constexpr (Bytes::kWords) {                       // four bytes per step: find the commas with SWAR compares
        for (uint32_t w0 = b & ~3u; w0 < e; w0 += 4u) {
            uint32_t z = swar_clip(swar_eq(s.word(w0), 0x2c2c2c2cu), w0, b, e);
            while (z) {
                const uint32_t i = w0 + ((uint32_t)(__ffs((int)z) - 1) >> 3);
                z &= z - 1u;
                if (nf < 6) { fb[nf] = st; fe[nf] = i; }
                nf++; st = i + 1;
            }
        }
        if (nf < 6) { fb[nf] = st; fe[nf] = e; }
        nf++;
    } else {
        for (uint32_t i = b; i <= e; i++) {
            if (i == e || s[i] == ',') {
                if (nf < 6) { fb[nf] = st; fe[nf] = i; }
                nf++; st = i + 1;
            }
        }
    }
    if (nf < 6) return RANK_SA_FIELDS;
    int64_t pos;
    if (!dev_parse_i64(s, fb[1], fe[1], &pos)) return RANK_SA_POS;
    const uint32_t sc = fe[2] - fb[2] == 1 ? s[fb[2]] : 0u;
    if (sc != '+' && sc != '-') return RANK_SA_STRAND;
    const uint32_t strand_neg = sc == '-';
    uint32_t sS = 0, sH = 0, sD = 0, sM = 0, sE = 0, sX = 0, ndig = 0;
    unsigned long long v = 0, key = 0; bool seenM = false;
    for (uint32_t i = fb[3]; i < fe[3]; i++) {
        const uint32_t c = s[i], d = c - '0';
        if (d <= 9u) { v = v * 10ull + d; if (v > 0x1ffffffffull) v = 0x1ffffffffull; ndig++; continue; }
        // op letters as bits of (c - '='):  = D H I M N P S X  ->  0 7 11 12 16 17 19 22 27
        const uint32_t x = c - '=';
        const bool isop = x < 28u && ((0x84B1881u >> x) & 1u);
        if (!isop || ndig == 0 || v > 0xffffffffull) return RANK_SA_CIGAR;
        const uint32_t n = (uint32_t)v;
        if (c == 'M') { sM += n; seenM = true; }
        else {
            if (c == 'S') sS += n; else if (c == 'D') sD += n; else if (c == 'H') sH += n;
            else if (c == '=') sE += n; else if (c == 'X') sX += n;
            if (!seenM && ((0x8401001u >> x) & 1u)) key += n;             // = I S X before the first M (utils.rs:33)
        }
        v = 0; ndig = 0;
    }
    if (!dev_parse_u8(s, fb[4], fe[4])) return RANK_SA_MAPQ;
    int64_t nm;
    if (!dev_parse_i64(s, fb[5], fe[5], &nm)) return RANK_SA_NM;
    uint32_t cb = fb[0];
    const uint32_t ce = fe[0];
    if (ce - cb >= 3 && s[cb] == 'c' && s[cb + 1] == 'h' && s[cb + 2] == 'r') cb += 3;
    out->chrom_ref = 0x80000000u | cb;
    out->chrom_len = ce - cb;
    out->start = (int64_t)((unsigned long long)pos - 1ull);
    out->end = out->start + (int64_t)sD + (int64_t)sM + (int64_t)sE + (int64_t)sX;
    out->key = (int64_t)key;
    out->clip_big = (sS > P.ins_clip_min || sH > P.ins_clip_min) ? 1u : 0u;
    out->strand_neg = strand_neg;
    return 0;
}

// Fast path of parse_supplementary_alignment (utils.rs:119-139) for the regular form every aligner writes: sa_parse_fast
// (exlr_sa_parse.cuh) on the staged bytes.  Anything else -- signs, longer numbers, missing or extra fields, odd bytes, trailing
// digits in the CIGAR -- returns false and the caller runs the exact dev_parse_piece above, which also yields the reference's
// panics.  For the accepted form both produce the same Seg: same field boundaries, u32-wrapping sums, u64 key.
__device__ __forceinline__ bool dev_parse_piece_fast(const uint8_t* p /* staged bytes */, uint32_t bias /* sa offset of p[0] */,
                                                     uint32_t b /* offsets into p */, uint32_t e, const DevParams& P, Seg* out, uint32_t m)
{
    SaFast f;
    if (!sa_parse_fast(p, b, e, &f, m)) return false;
    out->chrom_ref = 0x80000000u | (f.cb + bias);
    out->chrom_len = f.chrom_len;
    out->start = (int64_t)f.pos - 1;
    out->end = out->start + (int64_t)f.ref;
    out->key = (int64_t)f.key;
    out->clip_big = (f.clipS > P.ins_clip_min || f.clipH > P.ins_clip_min) ? 1u : 0u;
    out->strand_neg = f.strand_neg;
    return true;
}

template <class Bytes>
__device__ __forceinline__ uint32_t chrom_byte(const DevBatch& B, const Bytes& s, const Seg& g, uint32_t i)
{
    return (g.chrom_ref >> 31) ? s[(g.chrom_ref & 0x7fffffffu) + i] : (uint32_t)__ldg(B.ref_bytes + B.ref_off[g.chrom_ref] + i);
}

// String::cmp / as_bytes().cmp (utils.rs:76-77) on the "chr"-stripped names
template <class Bytes>
__device__ __forceinline__ int dev_chrom_cmp(const DevBatch& B, const Bytes& s, const Seg& a, const Seg& b)
{
    const uint32_t m = a.chrom_len < b.chrom_len ? a.chrom_len : b.chrom_len;
    for (uint32_t i = 0; i < m; i++) {
        const uint32_t x = chrom_byte(B, s, a, i), y = chrom_byte(B, s, b, i);
        if (x != y) return x < y ? -1 : 1;
    }
    return a.chrom_len < b.chrom_len ? -1 : (a.chrom_len > b.chrom_len ? 1 : 0);
}

// overlap (utils.rs:158-194); IEEE f64 divide and compare, like the Rust
__device__ __forceinline__ bool dev_overlap(int64_t as, int64_t ae, int64_t bs, int64_t be, double p)
{
    if (ae < bs || as > be) return false;
    const int64_t la = ae - as, lb = be - bs;
    const int64_t ml = la < lb ? la : lb;
    int64_t num;
    if (as < bs) num = ae < be ? ae - bs : be - bs;
    else         num = be < ae ? be - as : ae - as;
    const double ov = __ddiv_rn((double)num, (double)ml);
    return ov > p;
}

// Second half of a record's SA arm, shared by the fast and the fallback path: stable sort of the segments
// (main.rs:322), large-INS rules (main.rs:340-486), slot allocation, event emission (main.rs:488-516).
// Must be called by every lane of the warp (the slot allocation is warp-aggregated).
template <class Bytes>
__device__ __forceinline__ void k3b_finish(const DevBatch& B, const DevParams& P, const Bytes& s, uint32_t j, bool active,
                                           uint32_t r, bool dropped, Seg* segs, uint32_t nseg)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t n_ins = 0, ins_kind = 0;
    int64_t q1 = 0, q2 = 0, q0 = 0;
    if (active && nseg) {
        for (uint32_t i = 1; i < nseg; i++) {                             // stable insertion sort by key (main.rs:322)
            const Seg x = segs[i]; uint32_t k = i;
            while (k > 0 && segs[k - 1].key > x.key) { segs[k] = segs[k - 1]; k--; }
            segs[k] = x;
        }
        if (nseg == 2) {                                                  // main.rs:340-451
            const Seg& a = segs[0]; const Seg& b = segs[1];
            if (a.clip_big) {
                if (dev_chrom_cmp(B, s, a, b) == 0) {
                    if (a.strand_neg == b.strand_neg && dev_overlap(a.start, a.end, b.start, b.end, P.max_pct_overlap) && b.clip_big) {
                        int64_t q[4] = {a.start, a.end, b.start, b.end};
#pragma unroll
                        for (int x = 1; x < 4; x++) { const int64_t val = q[x]; int y = x; while (y > 0 && q[y - 1] > val) { q[y] = q[y - 1]; y--; } q[y] = val; }
                        q0 = q[0]; q1 = q[1]; q2 = q[2];
                        n_ins = 2; ins_kind = EXLR_KIND_INS_TWO_ALN;
                    }
                } else { n_ins = 1; ins_kind = EXLR_KIND_INS_ONE_ALN; }
            }
        } else if (nseg == 1) {                                           // main.rs:459-486
            if (segs[0].clip_big) { n_ins = 1; ins_kind = EXLR_KIND_INS_ONE_SEG; }
        }
    }
    // temp slots: one atomic per warp
    const uint32_t cnt = (active && nseg) ? n_ins + nseg - 1 : 0u;
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    uint32_t base = 0;
    const uint32_t wtotal = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 31 && wtotal) base = atomicAdd(&B.ctrl->n_saev, wtotal);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
    const uint32_t dm = __ballot_sync(0xffffffffu, active && dropped);
    if (lane == 0 && dm) atomicAdd(&B.ctrl->n_dropped, (uint32_t)__popc(dm));
    if (!active) return;
    B.sa_base[j] = base;
    B.csa[r] = dropped ? CSA_DROP : cnt;
    if (!cnt) return;
    if ((unsigned long long)base + cnt > B.max_events) { B.ctrl->overflow = 1; return; }
    exlr_event* dst = B.sa_ev + base;
    if (n_ins == 2) {
        const Seg& a = segs[0]; const Seg& b = segs[1];
        const uint32_t meta = EXLR_EV_META(1u, EXLR_KIND_INS_TWO_ALN, a.strand_neg, b.strand_neg);
        store_event(dst++, (int64_t)(uint32_t)q0, (int64_t)(uint32_t)q1, (int64_t)(uint32_t)q1, (int64_t)(uint32_t)q1, r, a.chrom_ref, b.chrom_ref, meta);
        store_event(dst++, (int64_t)(uint32_t)q0, (int64_t)(uint32_t)q2, (int64_t)(uint32_t)q2, (int64_t)(uint32_t)q2, r, a.chrom_ref, b.chrom_ref, meta);
    } else if (n_ins == 1) {
        const Seg& a = segs[0];
        const uint32_t meta = EXLR_EV_META(1u, ins_kind, a.strand_neg, a.strand_neg);
        store_event(dst++, (int64_t)(uint32_t)a.start, (int64_t)(uint32_t)a.end, (int64_t)(uint32_t)a.end, (int64_t)(uint32_t)a.end, r, a.chrom_ref, a.chrom_ref, meta);
    }
    for (uint32_t i = 1; i < nseg; i++) {                                 // main.rs:488-516
        const Seg* a = &segs[i - 1]; const Seg* b = &segs[i];
        int c = dev_chrom_cmp(B, s, *a, *b);                               // alignment_pos_cmp, utils.rs:75-86
        if (c == 0) c = a->start < b->start ? -1 : (a->start > b->start ? 1 : 0);
        if (c > 0) { const Seg* x = a; a = b; b = x; }
        store_event(dst++, a->start, a->end, b->start, b->end, r, a->chrom_ref, b->chrom_ref,
                    EXLR_EV_META(nseg - 1, EXLR_KIND_SPLIT, a->strand_neg, b->strand_neg));
    }
}

// the record's own alignment as segment 0 (main.rs:299-306)
__device__ __forceinline__ void seg_from_sums(const DevBatch& B, const DevParams& P, uint32_t r, uint32_t sS, uint32_t sH, int64_t refspan, int64_t ffm, Seg* out)
{
    const int32_t tid = B.tid[r];
    out->chrom_ref = (uint32_t)tid; out->chrom_len = B.ref_off[tid + 1] - B.ref_off[tid];
    out->start = (int64_t)B.pos[r]; out->end = out->start + refspan; out->key = ffm;
    out->clip_big = (sS > P.ins_clip_min || sH > P.ins_clip_min) ? 1u : 0u;
    out->strand_neg = (B.flag[r] & 0x10u) ? 1u : 0u;
}

__device__ __forceinline__ void k3b_record_seg(const DevBatch& B, const DevParams& P, uint32_t j, uint32_t r, Seg* out)
{
    const SaSum sum = B.sa_sum[j];
    seg_from_sums(B, P, r, sum.S, sum.H, sum.refspan, sum.ffm, out);
}

// The same segment straight from the record's CIGAR, walked by one thread (kernel 3a folded into kernel 3b: batches of short
// CIGARs, where a launch of its own costs more than the walk): per-op-type wrapping sums (main.rs:243-296), reference span
// D + M + = + X (split_read_event.rs:23-28), first-match offset (utils.rs:12-42).
__device__ __forceinline__ void k3b_own_seg(const DevBatch& B, const DevParams& P, uint32_t r, Seg* out)
{
    const unsigned long long o0 = B.cigar_off[r], o1 = B.cigar_off[r + 1];
    uint32_t sS = 0, sH = 0, sD = 0, sM = 0, sE = 0, sX = 0, bad = 0;
    unsigned long long ffm = 0; bool seenM = false;
    for (unsigned long long o = o0; o < o1; o += 4) {
        uint32_t vv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) vv[u] = o + u < o1 ? __ldg(B.cigar + o + u) : 0xfu;       // 0xf: not an op
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t v = vv[u], op = v & 15u, len = v >> 4;
            if (o + u < o1 && op > 8u) bad = 1;
            sM += op == 0u ? len : 0u; sD += op == 2u ? len : 0u; sS += op == 4u ? len : 0u;
            sH += op == 5u ? len : 0u; sE += op == 7u ? len : 0u; sX += op == 8u ? len : 0u;
            if (op == 0u) seenM = true;                                                       // utils.rs:28-29 stops at the first M
            if (!seenM && (op == 4u || op == 1u || op == 8u || op == 7u)) ffm += len;         // S I X =  (utils.rs:33)
        }
    }
    if (bad) report(B.ctrl, r, RANK_CIGAR_OP);
    seg_from_sums(B, P, r, sS, sH, (int64_t)sD + (int64_t)sM + (int64_t)sE + (int64_t)sX, (int64_t)ffm, out);
}

// Fallback: one SA record parsed start to finish by one thread (used when a tile's SA bytes or segment count do
// not fit the staged layout, e.g. -k far above the default).  Must be called by every lane of the warp.
template <bool FOLD, class Bytes>
__device__ __forceinline__ void k3b_record(const DevBatch& B, const DevParams& P, const Bytes& s, uint32_t j, bool active, Seg* local_segs)
{
    uint32_t r = 0, nseg = 0, err = 0;
    bool dropped = false;
    Seg* segs = local_segs;
    if (active) {
        r = B.sa_list[j];
        const uint32_t b0 = B.sa_off[r], e0 = B.sa_off[r + 1];
        const bool is_str = B.sa_kind[r] == EXLR_SA_STRING;
        unsigned long long pieces = 1;
        if (is_str) {
            for (uint32_t i = b0; i < e0; i++) pieces += s[i] == ';';
            if (pieces > P.max_supp_alignm) dropped = true;                 // main.rs:311-313: the whole record is skipped
        }
        if (!dropped && is_str && pieces + 1 > (unsigned long long)kLocalSegs) {
            const uint32_t need = (uint32_t)pieces + 1;
            const uint32_t at = atomicAdd(&B.ctrl->seg_pool_used, need);
            if ((unsigned long long)at + need <= B.seg_pool_cap) segs = B.seg_pool + at;
            else { B.ctrl->overflow = 1; dropped = true; }
        }
        if (!dropped) {
            if (FOLD) k3b_own_seg(B, P, r, &segs[0]); else k3b_record_seg(B, P, j, r, &segs[0]);
            nseg = 1;
            if (is_str) {
                uint32_t pb = b0;
                for (uint32_t i = b0; i <= e0 && !err; i++) {
                    if (i == e0 || s[i] == ';') {
                        if (i > pb) {                                         // filter(|x| x.len() > 0), main.rs:315
                            err = dev_parse_piece(s, pb, i, P, &segs[nseg]);
                            if (!err) nseg++;
                        }
                        pb = i + 1;
                    }
                }
            }
            if (!err && nseg - 1 >= (1u << 24)) err = RANK_SPLIT_COUNT;
            if (err) { report(B.ctrl, r, err); nseg = 0; }
        }
    }
    k3b_finish(B, P, s, j, active, r, dropped, segs, nseg);
}

// Fast path layout: the tile's SA bytes, a bit mask of their ';', the piece list and the segments all live in shared memory,
// and the work is re-flattened between phases so that the lanes of a warp always run the same code:
//   staging  thread per 16 bytes: copy to shared memory, one mask bit per byte that is ';'
//   phase 1  thread per record  : count ';' (the -k cap) and non-empty pieces from the mask     -> slots per record, block scan
//   phase 1b thread per record  : write the [begin,end) of every non-empty piece into the piece list
//   phase 2  thread per PIECE   : parse_supplementary_alignment from 8-byte windows (exlr_sa_parse.cuh) -> Seg
//   phase 3  thread per record  : record segment, sort, rules, emit
struct __align__(16) K3bSmem {
    uint8_t bytes[K3B_STAGE_BYTES + 16];               // (+16: the parser's windows read up to 11 bytes past a piece)
    uint32_t msemi[K3B_STAGE_BYTES / 32 + 4];          // bit i: staged byte i is ';'
    Seg segs[K3B_MAXP];                                // per record: its own alignment, then its pieces in SA order
    uint16_t pb[K3B_MAXP], pe[K3B_MAXP];               // piece list: [begin, end) as offsets into bytes[]
    uint16_t pslot[K3B_MAXP];                          // the piece's slot in segs[]
    uint8_t pread[K3B_MAXP];                           // the piece's record (index in the tile)
    uint32_t rerr[K3B_TILE];
    uint32_t wsum[K3B_THREADS / 32];
};
static_assert(offsetof(K3bSmem, segs) % 8 == 0, "Seg holds 64-bit fields");
// FOLD only (kept out of the plain layout: 2.6 KB more per CTA cost kernel 3b its seventh CTA per SM, and a grid sized for seven
// then ran a second, nearly empty wave -- 1.03 ms instead of 0.61 ms on configs[3])
struct __align__(16) K3bFoldSmem {
    SaSum sum[K3B_TILE];                               // what kernel 3a would have left in sa_sum, per record of the tile
    uint32_t rrec[K3B_TILE];                           // the tile's records
    uint8_t rlong[K3B_TILE];                           // the record's CIGAR is long: a whole warp walks it
};

// FOLD: kernel 3a's work is done here (the host folds it in for batches of short CIGARs): the thread that owns a record's slot
// in phase 2 walks the record's CIGAR while the other threads parse SA pieces.
template <bool FOLD>
__global__ void __launch_bounds__(K3B_THREADS, K3B_CTAS) k3b_sa_events(DevBatch B, DevParams P)
{
    extern __shared__ __align__(16) unsigned char k3b_smem_raw[];
    K3bSmem& S = *reinterpret_cast<K3bSmem*>(k3b_smem_raw);
    K3bFoldSmem& F = *reinterpret_cast<K3bFoldSmem*>(k3b_smem_raw + sizeof(K3bSmem));     // (only allocated for FOLD)
    // FOLD: the predecessor is kernel 0, whose list everything below reads.  Otherwise it is kernel 3a, whose CTAs trigger this
    // launch only after their own wait on kernel 0 has returned: kernel 0's list and count are complete and visible to any CTA
    // of this grid that runs, and kernel 3a's summaries are read in phase 3 only -- the wait is there, and staging, the piece
    // list and the parse run beside kernel 3a (which is bound by the latency of its CIGAR loads and leaves the issue slots idle).
    if (FOLD) griddep_wait();
    CtaTrace tr(B, 4);
    const uint32_t n_sa = B.ctrl->n_sa;
    const uint32_t n_tiles = (n_sa + K3B_TILE - 1) / K3B_TILE;
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
    Seg local_segs[kLocalSegs];
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t j0 = tile * K3B_TILE, j1 = min(j0 + (uint32_t)K3B_TILE, n_sa);
        const uint32_t span_b = B.sa_off[B.sa_list[j0]], span_e = B.sa_off[B.sa_list[j1 - 1] + 1];
        const uint32_t a0 = span_b & ~15u;
        const bool staged = span_e - a0 <= K3B_STAGE_BYTES;                   // block-uniform
        const uint32_t j = j0 + t;
        const bool active = t < K3B_TILE && j < j1;                           // the other threads only help with the pieces
        __syncthreads();                                                      // the previous tile's readers are done
        if (!staged) {
            GlobalBytes s{B.sa_bytes};
            griddep_wait();
            k3b_record<FOLD>(B, P, s, j, active, local_segs);
            continue;
        }
        for (uint32_t o = a0 + t * 16u; o < span_e; o += K3B_THREADS * 16u) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(B.sa_bytes + o));
            *reinterpret_cast<uint4*>(S.bytes + (o - a0)) = v;
            reinterpret_cast<uint16_t*>(S.msemi)[(o - a0) >> 4] =
                (uint16_t)(semi_bits(v.x) | (semi_bits(v.y) << 4) | (semi_bits(v.z) << 8) | (semi_bits(v.w) << 12));
        }
        if (FOLD) {
            // kernel 3a's work (main.rs:214-306, utils.rs:12-42): a pair of lanes walks the CIGAR of each of the tile's records --
            // its loads are in flight while the SA bytes are staged
            static_assert(K3B_THREADS >= 2 * K3B_TILE, "two lanes per record");
            const uint32_t k = t >> 1;
            const bool mine = t < 2 * K3B_TILE && j0 + k < j1;
            uint32_t rr = 0; unsigned long long o0 = 0, o1 = 0;
            if (mine) { rr = B.sa_list[j0 + k]; o0 = B.cigar_off[rr]; o1 = B.cigar_off[rr + 1]; }
            const bool is_long = mine && o1 - o0 > K3A_LONG;
            const uint32_t pm = __ballot_sync(0xffffffffu, mine && !is_long);
            if (mine && !is_long) {
                const K3aAcc a = k3a_walk<2>(B, rr, o0, o1, pm & (3u << (lane & ~1u)), lane & ~1u);
                if ((t & 1u) == 0u) { SaSum sm; sm.S = a.S; sm.H = a.H; sm.refspan = (int64_t)a.D + (int64_t)a.M + (int64_t)a.E + (int64_t)a.X; sm.ffm = (int64_t)a.ffm; sm.pad[0] = sm.pad[1] = 0; F.sum[k] = sm; }
            }
            if (mine && (t & 1u) == 0u) { F.rrec[k] = rr; F.rlong[k] = is_long ? 1 : 0; }
        }
        // phase 1: pieces per record
        uint32_t r = 0, b0 = 0, e0 = 0, slots = 0, npieces = 0;
        bool dropped = false, is_str = false;
        int32_t own_tid = 0, own_pos = 0; uint32_t own_flag = 0;              // the record's own alignment: asked for now, used in phase 3
        if (active) {
            r = B.sa_list[j]; b0 = B.sa_off[r]; e0 = B.sa_off[r + 1]; is_str = B.sa_kind[r] == EXLR_SA_STRING;
            own_tid = B.tid[r]; own_pos = B.pos[r]; own_flag = B.flag[r];
        }
        __syncthreads();
        SmemBytes s{S.bytes, a0};
        const uint32_t rb = b0 - a0, re = e0 - a0;                             // the record's string as offsets into bytes[]
        const bool scan = active && is_str && re > rb;
        const uint32_t k_first = rb >> 5, k_last = (re - 1u) >> 5;
        if (active) {
            // every ';' closes a piece, empty when it directly follows the previous ';' or the start of the string; the last
            // piece runs to the end of the string.  pieces = #';' + 1 (main.rs:309), non-empty ones get a slot (main.rs:315).
            uint32_t nsemi = 0, nonempty = 0;
            if (scan) {
                uint32_t carry = 1u;                                           // "the byte before is a ';' or the start"
                for (uint32_t k = k_first; k <= k_last; k++) {
                    uint32_t z = S.msemi[k];
                    if (k == k_first) z &= 0xffffffffu << (rb & 31u);
                    if (k == k_last) z &= 0xffffffffu >> (31u - ((re - 1u) & 31u));
                    uint32_t prev = (z << 1) | carry;
                    if (k == k_first) prev = (z << 1) | (1u << (rb & 31u));
                    nsemi += __popc(z);
                    nonempty += __popc(z & ~prev);
                    carry = z >> 31;
                }
                nonempty += ((S.msemi[k_last] >> ((re - 1u) & 31u)) & 1u) ^ 1u;   // the string does not end in ';': one more piece
            }
            if (is_str && (unsigned long long)nsemi + 1ull > P.max_supp_alignm) dropped = true;   // main.rs:311-313
            slots = dropped ? 0u : 1u + nonempty;
            npieces = dropped ? 0u : nonempty;
            S.rerr[t] = 0xffffffffu;
        }
        uint32_t incl = slots | (npieces << 16);                              // both running sums in one scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
        if (lane == 31) S.wsum[w] = incl;
        __syncthreads();
        uint32_t excl = incl - (slots | (npieces << 16)), total2 = 0;
#pragma unroll
        for (int k = 0; k < K3B_THREADS / 32; k++) { const uint32_t x = S.wsum[k]; if ((uint32_t)k < w) excl += x; total2 += x; }
        const uint32_t sb = excl & 0xffffu, total = total2 & 0xffffu, total_pieces = total2 >> 16;
        if (total > K3B_MAXP) {                                               // block-uniform: too many segments for the staged layout
            griddep_wait();
            k3b_record<FOLD>(B, P, s, j, active, local_segs);
            continue;
        }
        // phase 1b: piece list
        if (scan && !dropped) {
            uint32_t at = excl >> 16, slot = sb + 1u, pbeg = rb;
            auto piece_end = [&](uint32_t i) {
                if (i > pbeg) { S.pb[at] = (uint16_t)pbeg; S.pe[at] = (uint16_t)i; S.pslot[at] = (uint16_t)slot; S.pread[at] = (uint8_t)t; at++; slot++; }
                pbeg = i + 1u;
            };
            for (uint32_t k = k_first; k <= k_last; k++) {
                uint32_t z = S.msemi[k];
                if (k == k_first) z &= 0xffffffffu << (rb & 31u);
                if (k == k_last) z &= 0xffffffffu >> (31u - ((re - 1u) & 31u));
                while (z) { piece_end(k * 32u + (uint32_t)__ffs((int)z) - 1u); z &= z - 1u; }
            }
            if (re > pbeg) piece_end(re);
        }
        __syncthreads();
        // phase 2: one thread per piece
        for (uint32_t x0 = 0; x0 < total_pieces; x0 += K3B_THREADS) {        // block-uniform trip count
            const uint32_t x = x0 + t;
            const bool has = x < total_pieces;
            const uint32_t m = __ballot_sync(0xffffffffu, has);
            if (!has) continue;
            const uint32_t pbeg = S.pb[x], pend = S.pe[x], slot = S.pslot[x];
            if (dev_parse_piece_fast(S.bytes, a0, pbeg, pend, P, &S.segs[slot], m)) continue;
            const uint32_t err = dev_parse_piece(s, pbeg + a0, pend + a0, P, &S.segs[slot]);   // irregular piece: the exact parser decides
            if (err) atomicMin(&S.rerr[S.pread[x]], (x << 8) | err);          // the first failing piece in SA order wins
        }
        __syncthreads();
        if (FOLD) {                                                           // the odd long CIGAR of a short-CIGAR batch: a warp each
            for (uint32_t k = w; k < K3B_TILE; k += K3B_THREADS / 32) {
                if (j0 + k >= j1 || !F.rlong[k]) continue;                    // warp-uniform
                const uint32_t rr = F.rrec[k];
                const K3aAcc a = k3a_walk<32>(B, rr, B.cigar_off[rr], B.cigar_off[rr + 1], 0xffffffffu, 0);
                if (lane == 0) { SaSum sm; sm.S = a.S; sm.H = a.H; sm.refspan = (int64_t)a.D + (int64_t)a.M + (int64_t)a.E + (int64_t)a.X; sm.ffm = (int64_t)a.ffm; sm.pad[0] = sm.pad[1] = 0; F.sum[k] = sm; }
            }
            __syncthreads();
        }
        tr.mid();
        // phase 3: one thread per record
        uint32_t own_clen = 0;
        if (active && !dropped) own_clen = B.ref_off[own_tid + 1] - B.ref_off[own_tid];
        if (!FOLD) griddep_wait();                                            // kernel 3a's summaries
        uint32_t nseg = 0;
        if (active && !dropped) {
            uint32_t err = S.rerr[t];
            nseg = slots;
            if (err != 0xffffffffu) { report(B.ctrl, r, err & 0xffu); nseg = 0; }
            else {
                SaSum sm;
                if (!FOLD) sm = B.sa_sum[j]; else sm = F.sum[t];
                Seg& g = S.segs[sb];                                          // the record's own alignment as segment 0 (main.rs:299-306)
                g.chrom_ref = (uint32_t)own_tid; g.chrom_len = own_clen;
                g.start = (int64_t)own_pos; g.end = g.start + sm.refspan; g.key = sm.ffm;
                g.clip_big = (sm.S > P.ins_clip_min || sm.H > P.ins_clip_min) ? 1u : 0u;
                g.strand_neg = (own_flag & 0x10u) ? 1u : 0u;
            }
        }
        k3b_finish(B, P, s, j, active, r, dropped, &S.segs[sb < K3B_MAXP ? sb : 0], nseg);
    }
    tr.end();
}

// ======================================================================================
// launchers
// ======================================================================================
static int g_k3b_resident[2] = {K3B_CTAS, K3B_CTAS - 1};      // CTAs of kernel 3b that fit an SM at once: plain, FOLD (asked of the runtime below)

cudaError_t configure_sa_kernels()
{
    const int bytes[2] = {(int)sizeof(K3bSmem), (int)(sizeof(K3bSmem) + sizeof(K3bFoldSmem))};
    cudaError_t e = cudaFuncSetAttribute(k3b_sa_events<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes[0]);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k3b_sa_events<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes[1]);
    int n = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k3b_sa_events<false>, K3B_THREADS, bytes[0]);
    if (e == cudaSuccess && n > 0) g_k3b_resident[0] = n;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k3b_sa_events<true>, K3B_THREADS, bytes[1]);
    if (e == cudaSuccess && n > 0) g_k3b_resident[1] = n;
    return e;
}

// batches whose CIGARs are short get kernel 3a's work done inside kernel 3b (one launch less on the SA chain)
bool k3_fold(uint32_t mean_ops) { return mean_ops <= 64u; }

void launch_k3a(const DevBatch& B, const DevParams& P, uint32_t mean_ops, cudaStream_t st)
{
    // grid-stride over a device-side count: size for the worst case, cap at one resident wave (spare CTAs cost launch time).
    // Lanes per record by the batch's mean CIGAR length: the more records a warp walks side by side, the fewer waves.
    const uint32_t sms = (uint32_t)B.hc.sms;
    if (mean_ops <= 16) {
        const uint32_t g = min((B.n_reads + 127u) / 128u, sms * EXLR_RESIDENT_PER_SM(k3a_sa_cigar<2>, 256, 0));
        launch_dependent(k3a_sa_cigar<2>, g ? g : 1u, 256, 0, st, B, P);
    } else if (mean_ops <= 48) {
        const uint32_t g = min((B.n_reads + 63u) / 64u, sms * EXLR_RESIDENT_PER_SM(k3a_sa_cigar<4>, 256, 0));
        launch_dependent(k3a_sa_cigar<4>, g ? g : 1u, 256, 0, st, B, P);
    } else {
        const uint32_t g = min((B.n_reads + 31u) / 32u, sms * EXLR_RESIDENT_PER_SM(k3a_sa_cigar<8>, 256, 0));
        launch_dependent(k3a_sa_cigar<8>, g ? g : 1u, 256, 0, st, B, P);
    }
}

void launch_k3b(const DevBatch& B, const DevParams& P, bool fold, cudaStream_t st)
{
    // tiles of 64 SA records, grid-stride; the SA-record count lives on the device, so the grid is sized from the batch but
    // capped at the CTAs that are resident at once: spare CTAs of an over-sized grid cost a launch slot each just to read the
    // count and leave, and a CTA with a second tile doubles the kernel's span
    // (a grid larger than what is resident at once would run a second wave whose CTAs start when the first ones finish)
    const uint32_t gb = min((B.n_reads + K3B_TILE - 1u) / K3B_TILE, (uint32_t)B.hc.sms * (uint32_t)g_k3b_resident[fold ? 1 : 0]);
    if (fold) launch_dependent(k3b_sa_events<true>, gb ? gb : 1u, K3B_THREADS, sizeof(K3bSmem) + sizeof(K3bFoldSmem), st, B, P);
    else launch_dependent(k3b_sa_events<false>, gb ? gb : 1u, K3B_THREADS, sizeof(K3bSmem), st, B, P);
}

}  // namespace exlr
