// exlr_sa_parse.cuh — one regular SA piece parsed from 8-byte windows (reference src/utils.rs:88-139, 12-42;
// src/split_read_event.rs:23-28).
//
// The regular form is what every aligner writes:
//     chrom,<1-9 digits>,<+|->,(<1-9 digits><op>)+,<1-3 digits <= 255>,<1-9 digits>
// Kernel 3b (exlr_sa.cu) calls sa_parse_fast on a piece staged in shared memory; anything the function does not accept is
// parsed again by the exact byte-by-byte parser, which also yields the reference's panics.  For an accepted piece both give the
// same fields.  Numbers are not walked digit by digit: the eight bytes at a number's first digit are loaded as two words, the
// length of the digit run comes from one SWAR compare, and the value from three multiplies -- every lane of a warp runs the same
// few instructions per field, whatever the field's length.
//
// The header compiles for the host as well (tests/sa_fast_harness.cpp fuzzes it against a plain restatement of the reference
// parser on the CPU); nothing in the product calls the host build.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define EXLR_HD __host__ __device__ __forceinline__
#else
#define EXLR_HD inline
#endif

namespace exlr {

EXLR_HD uint32_t sa_ffs(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)x);
#else
    return (uint32_t)__builtin_ffs((int)x);
#endif
}

// low 32 bits of (hi:lo) >> sh, sh in 0..31
EXLR_HD uint32_t sa_fshr(uint32_t lo, uint32_t hi, uint32_t sh)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
}

// the 8 bytes at byte offset s of a 4-byte aligned buffer (little endian: byte s is bits 0..7 of lo); reads the three aligned
// words that cover them, so the buffer must be readable up to 11 bytes past s
struct SaWin { uint32_t lo, hi; };
EXLR_HD SaWin sa_window(const uint32_t* words, uint32_t s)
{
    const uint32_t k = s >> 2, sh = (s & 3u) * 8u;
    const uint32_t w0 = words[k], w1 = words[k + 1], w2 = words[k + 2];
    SaWin w; w.lo = sa_fshr(w0, w1, sh); w.hi = sa_fshr(w1, w2, sh);
    return w;
}

// 0x80 in every byte of x that equals the byte replicated in c4 (exact: no carries cross byte lanes)
EXLR_HD uint32_t sa_eq(uint32_t x, uint32_t c4)
{
    const uint32_t y = x ^ c4;
    return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y | 0x7f7f7f7fu);
}
// 0x80 in every byte of x that is not '0'..'9'
EXLR_HD uint32_t sa_nondigit(uint32_t x)
{
    const uint32_t y = x ^ 0x30303030u;                     // digits -> 0..9
    return (((y & 0x7f7f7f7fu) + 0x76767676u) | y) & 0x80808080u;
}
// index (0..3) of the lowest flagged byte of a nonzero SWAR flag word
EXLR_HD uint32_t sa_first(uint32_t z) { return (sa_ffs(z) >> 3) - 1u; }

// length of the digit run at the start of the window: 0..8
EXLR_HD uint32_t sa_digits(SaWin w)
{
    const uint32_t f0 = sa_nondigit(w.lo), f1 = sa_nondigit(w.hi);
    return f0 ? sa_first(f0) : (f1 ? sa_first(f1) + 4u : 8u);
}
// byte k (0..7) of the window
EXLR_HD uint32_t sa_byte(SaWin w, uint32_t k)
{
    return (uint32_t)(((((uint64_t)w.hi) << 32) | w.lo) >> (8u * k)) & 0xffu;
}
// value of the first L (1..8) bytes of the window read as decimal digits
EXLR_HD uint32_t sa_value(SaWin w, uint32_t L)
{
    // the run moves to the top of the eight bytes (the low bytes become leading zeros); then pairs, fours, all eight
    const uint64_t W = ((((uint64_t)w.hi) << 32) | w.lo) << (8u * (8u - L));
    uint32_t a = (uint32_t)W & 0x0f0f0f0fu, b = (uint32_t)(W >> 32) & 0x0f0f0f0fu;
    a = ((a * 2561u) >> 8) & 0x00ff00ffu; a = (a * 6553601u) >> 16;
    b = ((b * 2561u) >> 8) & 0x00ff00ffu; b = (b * 6553601u) >> 16;
    return a * 10000u + b;
}

// what the piece says, before it becomes a Seg (exlr_device.cuh)
struct SaFast {
    uint32_t cb, chrom_len;      // chrom name without a leading "chr": offset (same base as b, e) and length
    uint32_t pos;                // 1-based POS as written
    uint32_t ref;                // = + D + M + X     (split_read_event.rs:23-28)
    uint32_t key;                // = + I + S + X before the first M   (utils.rs:12-42)
    uint32_t clipS, clipH;       // sums of S and of H
    uint32_t strand_neg;
};

#if defined(__CUDA_ARCH__)
#define EXLR_REJOIN(m) __syncwarp(m)
#else
#define EXLR_REJOIN(m) ((void)(m))
#endif

// p: 4-byte aligned buffer, readable up to p + e + 192 (bytes past e may hold anything); [b, e): the piece.
// m: the lanes of the warp that call this together (device only; they re-join after each loop).
// Returns false for anything but the regular form.  No early exits: a failed check only clears `ok`, so the lanes stay together.
EXLR_HD bool sa_parse_fast(const uint8_t* p, uint32_t b, uint32_t e, SaFast* o, uint32_t m)
{
    const uint32_t* words = reinterpret_cast<const uint32_t*>(p);
    // chrom: up to the first ','
    SaWin w = sa_window(words, b);
    const bool chr = (w.lo & 0x00ffffffu) == 0x00726863u;                       // "chr"
    uint32_t i = b, c0 = e;
    for (;;) {
        const uint32_t z0 = sa_eq(w.lo, 0x2c2c2c2cu), z1 = sa_eq(w.hi, 0x2c2c2c2cu);
        if (z0 | z1) { c0 = i + (z0 ? sa_first(z0) : sa_first(z1) + 4u); break; }
        i += 8u;
        if (i >= e) break;
        w = sa_window(words, i);
    }
    EXLR_REJOIN(m);
    bool ok = c0 < e;
    c0 = c0 < e ? c0 : b;                                                       // (keeps the reads below near the piece)
    // pos: 1..9 digits and ','
    uint32_t s = c0 + 1u;
    w = sa_window(words, s);
    uint32_t L = sa_digits(w);
    uint32_t pos = sa_value(w, L ? L : 1u);
    if (L == 8u) { const uint32_t d = (uint32_t)p[s + 8u] - '0'; if (d <= 9u) { pos = pos * 10u + d; L = 9u; } }
    ok &= L != 0u;
    const uint32_t q = s + L;
    const uint32_t t0 = p[q], sc = p[q + 1u], t2 = p[q + 2u];                    // ',' strand ','
    ok &= t0 == ',' && (sc == '+' || sc == '-') && t2 == ',' && q + 2u < e;
    s = q + 3u;
    // CIGAR text (utils.rs:88-117, 12-42): one window per op.  With at most 15 ops of less than 2^28 each no sum can wrap a u32,
    // so one accumulator serves = + D + M + X; longer CIGARs or lengths go to the exact parser.
    //   = D H I M N P S X  ->  c - '=' = 0 7 11 12 16 17 19 22 27
    uint32_t ref = 0, sS = 0, sH = 0, key = 0, seenM = 0, nops = 0, bad = 0, big = 0, closed = 0;
    for (uint32_t it = 0; it < 16u; it++) {
        w = sa_window(words, s);
        L = sa_digits(w);
        uint32_t n = sa_value(w, L ? L : 1u);
        uint32_t c;
        if (L < 8u) c = sa_byte(w, L);
        else {
            c = p[s + 8u];
            const uint32_t d = c - '0';
            if (d <= 9u) { n = n * 10u + d; L = 9u; c = p[s + 9u]; }
        }
        if (c == ',') { bad |= L; closed = 1u; break; }                         // the closing comma follows an op letter directly
        const uint32_t x = c - '=', xs = x & 31u, in = x < 28u ? 1u : 0u;        // `in` voids the class bits of bytes beyond 'X'
        bad |= (((0x84B1881u >> xs) & in) ^ 1u) | (L == 0u ? 1u : 0u);
        big |= n;
        ref += n * ((0x8010081u >> xs) & in);                                   // = D M X
        sS += n * ((0x0400000u >> xs) & in);
        sH += n * ((0x0000800u >> xs) & in);
        key += n * ((0x8401001u >> xs) & in & (seenM ^ 1u));                    // = I S X before the first M (utils.rs:33)
        seenM |= (0x0010000u >> xs) & in;
        s += L + 1u; nops++;
    }
    EXLR_REJOIN(m);
    ok &= closed && !bad && nops - 1u < 15u && (big >> 28) == 0u && s < e;      // s: the closing comma
    // mapq: 1..3 digits <= 255 and ','
    s += 1u;
    w = sa_window(words, s);
    L = sa_digits(w);
    const uint32_t mq = sa_value(w, L ? L : 1u);
    ok &= L - 1u < 3u && mq <= 255u && sa_byte(w, L < 8u ? L : 0u) == ',' && s + L < e;
    // NM: 1..9 digits up to the end of the piece (the value is not used, utils.rs:135)
    s += L + 1u;
    w = sa_window(words, s);
    L = sa_digits(w);
    if (L == 8u && (uint32_t)p[s + 8u] - '0' <= 9u) L = 9u;
    const uint32_t rem = e - s;
    ok &= s < e && rem <= 9u && L >= rem;
    const uint32_t skip = (chr && c0 - b >= 3u) ? 3u : 0u;
    o->cb = b + skip; o->chrom_len = c0 - b - skip;
    o->pos = pos; o->ref = ref; o->key = key; o->clipS = sS; o->clipH = sH; o->strand_neg = sc == '-';
    return ok;
}

}  // namespace exlr
