// TEST INFRASTRUCTURE (host build of csrc/exlr_sa_parse.cuh): fuzzes the window-based SA piece parser of kernel 3b against a
// plain byte-by-byte restatement of parse_supplementary_alignment + parse_cigar + find_first_match_pos
// (reference src/utils.rs:12-42, 88-139; src/split_read_event.rs:23-28).
// Property checked: whenever the fast parser accepts a piece, the plain parser accepts it too and both give the same fields.
// usage: sa_fast_harness <seed> <cases>      exit 0 = ok; prints the first failing piece otherwise
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../excord_lr_b200/csrc/exlr_sa_parse.cuh"

struct Plain { bool ok; std::string chrom; int64_t start, end, key; uint32_t S, H; bool neg; };

static bool parse_i64(const std::string& t, int64_t* out)
{
    size_t i = 0; bool neg = false;
    if (t.empty()) return false;
    if (t[0] == '+' || t[0] == '-') { neg = t[0] == '-'; i = 1; }
    if (i == t.size()) return false;
    unsigned long long v = 0;
    for (; i < t.size(); i++) {
        if (t[i] < '0' || t[i] > '9') return false;
        if (v > 1000000000000000000ull) return false;          // (far beyond anything the fast form accepts)
        v = v * 10 + (unsigned)(t[i] - '0');
    }
    *out = neg ? -(int64_t)v : (int64_t)v;
    return true;
}

static Plain plain_parse(const std::string& piece)
{
    Plain r{}; r.ok = false;
    std::vector<std::string> f; size_t st = 0;
    for (size_t i = 0; i <= piece.size(); i++) if (i == piece.size() || piece[i] == ',') { f.push_back(piece.substr(st, i - st)); st = i + 1; }
    if (f.size() < 6) return r;
    int64_t pos; if (!parse_i64(f[1], &pos)) return r;
    if (f[2] != "+" && f[2] != "-") return r;
    uint32_t sS = 0, sH = 0, sD = 0, sM = 0, sE = 0, sX = 0; unsigned long long key = 0, v = 0; int nd = 0; bool seenM = false;
    for (char ch : f[3]) {
        if (ch >= '0' && ch <= '9') { v = v * 10 + (unsigned)(ch - '0'); if (v > 0x1ffffffffull) v = 0x1ffffffffull; nd++; continue; }
        if (!strchr("=DHIMNPSX", ch) || nd == 0 || v > 0xffffffffull) return r;
        const uint32_t n = (uint32_t)v;
        if (ch == 'M') { sM += n; seenM = true; }
        else {
            if (ch == 'S') sS += n; else if (ch == 'D') sD += n; else if (ch == 'H') sH += n; else if (ch == '=') sE += n; else if (ch == 'X') sX += n;
            if (!seenM && (ch == '=' || ch == 'I' || ch == 'S' || ch == 'X')) key += n;
        }
        v = 0; nd = 0;
    }
    {   // mapq: u8
        const std::string& t = f[4]; size_t i = 0; if (t.empty()) return r; if (t[0] == '+') i = 1; if (i == t.size()) return r;
        unsigned q = 0; for (; i < t.size(); i++) { if (t[i] < '0' || t[i] > '9') return r; q = q * 10 + (unsigned)(t[i] - '0'); if (q > 255) return r; }
    }
    int64_t nm; if (!parse_i64(f[5], &nm)) return r;
    r.chrom = f[0].compare(0, 3, "chr") == 0 ? f[0].substr(3) : f[0];
    r.start = pos - 1; r.end = r.start + (int64_t)sD + (int64_t)sM + (int64_t)sE + (int64_t)sX; r.key = (int64_t)key;
    r.S = sS; r.H = sH; r.neg = f[2] == "-"; r.ok = true;
    return r;
}

static std::mt19937_64 rng;
static uint32_t rnd(uint32_t n) { return (uint32_t)(rng() % n); }

static std::string digits(int n, bool nonzero_first)
{
    std::string s;
    for (int i = 0; i < n; i++) s += (char)('0' + ((i == 0 && nonzero_first) ? 1 + rnd(9) : rnd(10)));
    return s;
}

static std::string regular_piece()
{
    static const char* chroms[] = {"chr1", "chr20", "chrX", "20", "1", "chrUn_KI270742v1", "HLA-A*01:01:01:01", "chr", "ch", "c", "", "chrM", "chr1_KI270706v1_random", "7"};
    std::string s = chroms[rnd(14)];
    s += ","; s += digits(1 + rnd(rnd(4) ? 9 : 11), rnd(8) != 0);
    s += ","; s += rnd(2) ? "+" : "-"; s += ",";
    const int nops = rnd(16) ? 1 + rnd(6) : rnd(20);
    for (int k = 0; k < nops; k++) { s += digits(rnd(12) ? 1 + rnd(5) : 1 + rnd(10), rnd(6) != 0); s += "=DHIMNPSX"[rnd(9)]; }
    s += ","; s += rnd(10) ? std::to_string(rnd(61)) : digits(1 + rnd(4), false);
    s += ","; s += digits(1 + rnd(rnd(5) ? 4 : 10), false);
    return s;
}

static std::string mutate(std::string s)
{
    static const char junk[] = "0123456789,,;+-MIDNSHP=XZ chr\t*:_.";
    const int n = 1 + rnd(3);
    for (int k = 0; k < n && !s.empty(); k++) {
        const uint32_t at = rnd((uint32_t)s.size());
        switch (rnd(4)) {
        case 0: s[at] = junk[rnd(sizeof(junk) - 1)]; break;
        case 1: s.erase(at, 1 + rnd(3)); break;
        case 2: s.insert(at, 1, junk[rnd(sizeof(junk) - 1)]); break;
        default: s.insert(at, digits(1 + rnd(9), false)); break;
        }
    }
    for (char& c : s) if (c == ';') c = ',';                     // a piece never holds ';' (the caller splits on it)
    return s;
}

int main(int argc, char** argv)
{
    const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    const long cases = argc > 2 ? atol(argv[2]) : 200000;
    rng.seed(seed);
    long accepted = 0, plain_ok = 0;
    std::vector<uint8_t> buf(4096 + 256);
    for (long it = 0; it < cases; it++) {
        std::string piece = regular_piece();
        if (rnd(3) == 0) piece = mutate(piece);
        // the piece sits at a random byte offset; what follows it is another piece, digits, or noise
        const uint32_t b = 16 + rnd(40);
        const uint32_t e = b + (uint32_t)piece.size();
        for (uint32_t k = 0; k < e + 224 && k < buf.size(); k++) if (k < b || k >= e) buf[k] = (uint8_t)(rnd(4) ? "0123456789"[rnd(10)] : (rnd(2) ? "MSc,;h"[rnd(6)] : (uint8_t)rnd(256)));
        memcpy(buf.data() + b, piece.data(), piece.size());
        if (rnd(2)) buf[e] = ';';
        exlr::SaFast f{};
        const bool ok = exlr::sa_parse_fast(buf.data(), b, e, &f, 0xffffffffu);
        const Plain r = plain_parse(piece);
        plain_ok += r.ok;
        if (!ok) continue;
        accepted++;
        const std::string chrom((const char*)buf.data() + f.cb, f.chrom_len);
        const bool same = r.ok && chrom == r.chrom && (int64_t)f.pos - 1 == r.start && r.start + (int64_t)f.ref == r.end && (int64_t)f.key == r.key &&
                          f.clipS == r.S && f.clipH == r.H && (f.strand_neg != 0) == r.neg;
        if (!same) {
            printf("MISMATCH seed %llu case %ld piece \"%s\" (offset %u): fast chrom \"%s\" pos %u ref %u key %u S %u H %u neg %u; plain ok %d chrom \"%s\" start %lld end %lld key %lld S %u H %u neg %d\n",
                   (unsigned long long)seed, it, piece.c_str(), b, chrom.c_str(), f.pos, f.ref, f.key, f.clipS, f.clipH, f.strand_neg, (int)r.ok, r.chrom.c_str(),
                   (long long)r.start, (long long)r.end, (long long)r.key, r.S, r.H, (int)r.neg);
            return 1;
        }
    }
    printf("ok: %ld cases, %ld parse in the reference, %ld taken by the fast parser\n", cases, plain_ok, accepted);
    // the regular form must actually be taken (or kernel 3b would silently run its slow path)
    return accepted * 2 > plain_ok ? 0 : 2;
}
