// exlr_synth.cpp — deterministic synthetic long-read alignment batches.
//
// Produces the structure-of-arrays batch of include/exlr.h directly (what the host packer
// would produce from a BAM), for the five BASELINE.json configs:
//   profile 0  HiFi    c1/c2/c5: N(15 kb, 2 kb) reads, M-style CIGARs with 1-3 bp indels at
//              0.1 %/bp, injected SV sites (DEL/INS 50-5000 bp, DEL pairs with gaps 0-8, DEL
//              triples 2 bp apart, sizes 49/50/51), ~5 % split molecules with reciprocal SA
//              tags (primary S-clipped, supplementary H-clipped), large-INS clip cases.
//   profile 1  ONT     c3: log-normal lengths (N50 ~50 kb, cap 1 Mb), 5 % indel error, 80 %
//              M-style / 20 % =/X-style CIGARs, `n_ultra` reads with > 65 535 ops.
//   profile 2  SPLIT   c4: every molecule split into 2-9 alignments (SA with 1-8 segments,
//              geometric), short CIGARs.
// Records are coordinate sorted (tid, pos; unmapped last) like a sorted BAM.  Everything is a
// pure function of (profile, seed, n_molecules, flags): the same bytes here and on the GPU box.
//
// Not part of the hot path and not the oracle: bench.py / tests use it to make inputs.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

const char* kNames[25] = {"chr1","chr2","chr3","chr4","chr5","chr6","chr7","chr8","chr9","chr10","chr11","chr12",
    "chr13","chr14","chr15","chr16","chr17","chr18","chr19","chr20","chr21","chr22","chrX","chrY","chrM"};
// GRCh38 primary assembly lengths
const int64_t kLens[25] = {248956422,242193529,198295559,190214555,181538259,170805979,159345973,145138636,138394717,
    133797422,135086622,133275309,114364328,107043718,101991189,90338345,83257441,80373285,58617616,64444167,46709983,
    50818468,156040895,57227415,16569};

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
                      z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }          // [0,1)
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
    int64_t range(int64_t lo, int64_t hi) { return lo + (int64_t)below((uint64_t)(hi - lo + 1)); }  // inclusive
    double normal() { double u = uni(), v = uni(); if (u < 1e-300) u = 1e-300; return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v); }
    int64_t geometric(double mean) { double u = uni(); if (u < 1e-300) u = 1e-300; return 1 + (int64_t)(-std::log(u) * (mean - 1.0 < 0.0 ? 0.0 : mean - 1.0)); }
};

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };
inline uint32_t cg(uint32_t len, uint32_t op) { return (len << 4) | op; }

struct Rec {
    int32_t tid, pos; uint16_t flag; uint8_t mapq, sa_kind;
    uint64_t cig_begin, cig_end;    // into flat cigar store
    uint64_t sa_begin, sa_end;      // into flat sa store
    uint64_t qid;
};

struct SvSite { int tid; int64_t pos; int kind; uint32_t len[3]; uint32_t gap[2]; };  // kind 0 DEL,1 INS,2 DEL pair,3 DEL triple

struct Gen {
    int profile; uint64_t seed; bool chr20_only;
    double sa_frac;
    std::vector<uint32_t> cigar; std::vector<uint8_t> sa; std::vector<Rec> recs;
    std::vector<SvSite> sites;
    Rng rng;
    Gen(int p, uint64_t s, bool c20) : profile(p), seed(s), chr20_only(c20), sa_frac(0.05), rng(s * 0x2545F4914F6CDD1Dull + 12345) {}

    int pick_tid() {
        if (chr20_only) return 19;
        // weight by length over chr1..chrY (chrM excluded)
        static double cum[24]; static bool init = false;
        if (!init) { double t = 0; for (int i = 0; i < 24; i++) { t += (double)kLens[i]; cum[i] = t; } init = true; }
        double x = rng.uni() * cum[23];
        int i = 0; while (i < 23 && x >= cum[i]) i++;
        return i;
    }

    // alignment body over `qlen` query bases; appends ops, returns reference span.  style 0 = M, 1 = =/X.
    // sv != nullptr injects that site's event(s) once the walk passes `sv_at` query bases.
    int64_t body(int64_t qlen, double indel_rate, int style, const SvSite* sv, int64_t sv_at) {
        int64_t q = 0, ref = 0; bool sv_done = sv == nullptr;
        double mean_run = 1.0 / indel_rate;
        while (q < qlen) {
            int64_t run = rng.geometric(mean_run);
            if (run > qlen - q) run = qlen - q;
            if (!sv_done && q + run >= sv_at) {
                int64_t first = sv_at - q; if (first < 1) first = 1; if (first > run) first = run;
                match(first, style); q += first; ref += first;
                ref += inject(*sv, style);
                sv_done = true;
                continue;
            }
            match(run, style); q += run; ref += run;
            if (q >= qlen) break;
            uint32_t n = (uint32_t)rng.range(1, 3);
            if (rng.next() & 1) { cigar.push_back(cg(n, OP_D)); ref += n; }
            else { if ((int64_t)n > qlen - q) n = (uint32_t)(qlen - q); cigar.push_back(cg(n, OP_I)); q += n; }
        }
        return ref;
    }
    void match(int64_t run, int style) {
        if (style == 0) { cigar.push_back(cg((uint32_t)run, OP_M)); return; }
        // =/X style: mismatches at ~1 %
        int64_t left = run;
        while (left > 0) {
            int64_t eq = rng.geometric(100.0); if (eq > left) eq = left;
            cigar.push_back(cg((uint32_t)eq, OP_EQ)); left -= eq;
            if (left > 0) { cigar.push_back(cg(1, OP_X)); left -= 1; }
        }
    }
    int64_t inject(const SvSite& s, int style) {
        int64_t ref = 0;
        switch (s.kind) {
        case 0: cigar.push_back(cg(s.len[0], OP_D)); ref += s.len[0]; break;
        case 1: cigar.push_back(cg(s.len[0], OP_I)); break;
        case 2: cigar.push_back(cg(s.len[0], OP_D)); ref += s.len[0];
                if (s.gap[0]) { match(s.gap[0], style); ref += s.gap[0]; }
                cigar.push_back(cg(s.len[1], OP_D)); ref += s.len[1]; break;
        default:
                for (int k = 0; k < 3; k++) {
                    cigar.push_back(cg(s.len[k], OP_D)); ref += s.len[k];
                    if (k < 2) { match(s.gap[k] ? s.gap[k] : 2, style); ref += s.gap[k] ? s.gap[k] : 2; }
                }
        }
        return ref;
    }

    void make_sites(int n_each) {
        Rng r(seed ^ 0xabcdef12345ull);
        auto lognuni = [&](double lo, double hi) { return (uint32_t)std::llround(std::exp(std::log(lo) + r.uni() * (std::log(hi) - std::log(lo)))); };
        auto place = [&](SvSite& s) { s.tid = chr20_only ? 19 : (int)r.below(24); s.pos = 1000000 + (int64_t)r.below((uint64_t)(kLens[s.tid] - 3000000)); };
        const uint32_t edge[3] = {49, 50, 51};
        for (int i = 0; i < n_each; i++) { SvSite s{}; place(s); s.kind = 0; s.len[0] = i < 3 ? edge[i] : lognuni(50, 5000); sites.push_back(s); }
        for (int i = 0; i < n_each; i++) { SvSite s{}; place(s); s.kind = 1; s.len[0] = i < 3 ? edge[i] : lognuni(50, 5000); sites.push_back(s); }
        for (int i = 0; i < n_each / 2; i++) { SvSite s{}; place(s); s.kind = 2; s.len[0] = lognuni(50, 500); s.len[1] = lognuni(50, 500); s.gap[0] = (uint32_t)(i % 9); sites.push_back(s); }
        for (int i = 0; i < n_each / 4; i++) { SvSite s{}; place(s); s.kind = 3; for (int k = 0; k < 3; k++) s.len[k] = 50 + (uint32_t)r.below(10); s.gap[0] = s.gap[1] = 2; sites.push_back(s); }
    }

    uint16_t noise_flags(uint8_t* mapq) {
        uint16_t f = (rng.next() & 1) ? 0x10 : 0;
        double u = rng.uni();
        if (u < 0.01) f |= 0x100; else if (u < 0.015) f |= 0x400;
        *mapq = rng.uni() < 0.02 ? 0 : 60;
        return f;
    }

    void push_rec(int32_t tid, int32_t pos, uint16_t flag, uint8_t mapq, uint64_t cb, uint64_t qid) {
        Rec r{}; r.tid = tid; r.pos = pos; r.flag = flag; r.mapq = mapq; r.sa_kind = 0;
        r.cig_begin = cb; r.cig_end = cigar.size(); r.sa_begin = r.sa_end = sa.size(); r.qid = qid;
        recs.push_back(r);
    }

    int64_t draw_len() {
        if (profile == 1) {                       // ONT: log-normal, N50 ~ 50 kb, cap 1 Mb
            double l = std::exp(std::log(22000.0) + 0.9 * rng.normal());
            l = std::min(std::max(l, 1000.0), 1000000.0);
            return (int64_t)l;
        }
        if (profile == 2) return rng.range(2000, 12000);
        double l = 15000.0 + 2000.0 * rng.normal();
        l = std::min(std::max(l, 5000.0), 25000.0);
        return (int64_t)l;
    }

    // one linear (unsplit) molecule
    void linear(uint64_t qid, int64_t force_len) {
        uint8_t mapq; uint16_t flag = noise_flags(&mapq);
        if (rng.uni() < 0.002) {                   // unmapped: tid -1, no CIGAR (filtered by -F 4)
            push_rec(-1, -1, (uint16_t)(flag | 0x4), 0, cigar.size(), qid); return;
        }
        int64_t len = force_len ? force_len : draw_len();
        int style = (profile == 1 && rng.uni() < 0.2) ? 1 : 0;
        double rate = profile == 1 ? 0.05 : 0.001;
        const SvSite* sv = nullptr; int64_t sv_at = 0; int tid; int64_t pos;
        if (!sites.empty() && rng.uni() < 0.12) {
            sv = &sites[rng.below(sites.size())];
            sv_at = rng.range(500, len - 500);
            tid = sv->tid; pos = sv->pos - sv_at; if (pos < 0) pos = 0;
        } else {
            tid = pick_tid(); pos = (int64_t)rng.below((uint64_t)(kLens[tid] - len - 20000));
        }
        uint64_t cb = cigar.size();
        int64_t lclip = rng.uni() < 0.3 ? rng.range(1, 200) : 0, rclip = rng.uni() < 0.3 ? rng.range(1, 200) : 0;
        if (lclip) cigar.push_back(cg((uint32_t)lclip, OP_S));
        body(len - lclip - rclip, rate, style, sv, sv_at);
        if (rclip) cigar.push_back(cg((uint32_t)rclip, OP_S));
        push_rec(tid, (int32_t)pos, flag, mapq, cb, qid);
    }

    struct Piece { int tid; int64_t pos; int strand; int64_t q0, q1; int64_t refspan; uint64_t cb, ce; uint32_t nm; };

    // a split molecule: `nseg` alignments of consecutive query intervals; one primary + supplementaries with reciprocal SA
    void split(uint64_t qid, int nseg, int mode /*0 mixed,1 large-ins same chrom overlapping*/, int64_t clip_target) {
        int64_t len = draw_len();
        if (mode == 1) len = 20000 + clip_target;
        if (len < 400LL * nseg) len = 400LL * nseg;
        std::vector<int64_t> cuts(nseg + 1); cuts[0] = 0; cuts[nseg] = len;
        for (int i = 1; i < nseg; i++) cuts[i] = rng.range(150, len - 150);
        std::sort(cuts.begin(), cuts.end());
        for (int i = 1; i <= nseg; i++) if (cuts[i] <= cuts[i - 1] + 100) cuts[i] = cuts[i - 1] + 100;
        len = cuts[nseg];
        if (mode == 1) { cuts[1] = len - clip_target; }
        std::vector<Piece> ps(nseg);
        int tid0 = pick_tid(); int64_t base = 2000000 + (int64_t)rng.below((uint64_t)(kLens[tid0] - 6000000));
        int strand0 = (rng.next() & 1) ? -1 : 1;
        int64_t last0 = base;                       // end of the last piece placed on tid0
        for (int i = 0; i < nseg; i++) {
            Piece& p = ps[i]; p.q0 = cuts[i]; p.q1 = cuts[i + 1];
            double u = rng.uni();
            if (mode == 1) { p.tid = tid0; p.strand = strand0; p.pos = i == 0 ? base : base + (ps[0].q1 - ps[0].q0) - rng.range(1, 300); }
            else if (i == 0) { p.tid = tid0; p.strand = strand0; p.pos = base; }
            else if (u < 0.5) { p.tid = tid0; p.strand = strand0; p.pos = last0 + rng.range(50, 200000); }                                         // deletion-like
            else if (u < 0.7) { p.tid = tid0; p.strand = -strand0; p.pos = base + rng.range(-500000, 500000); }                                    // inversion
            else if (u < 0.8) { p.tid = tid0; p.strand = strand0; p.pos = last0 - rng.range(0, 5000); }                                            // dup-like / overlapping
            else { p.tid = chr20_only ? (int)rng.below(24) : pick_tid(); p.strand = (rng.next() & 1) ? -1 : 1;
                   p.pos = 1000000 + (int64_t)rng.below((uint64_t)(kLens[p.tid] - 3000000)); }                                                 // inter-chromosomal
            if (p.pos < 0) p.pos = 0;
            if (p.pos > kLens[p.tid] - 2000000) p.pos = kLens[p.tid] - 2000000;
            if (p.tid == tid0) last0 = p.pos + (p.q1 - p.q0);
        }
        // records: piece `prim` is the primary (soft clips), others supplementary (hard clips, 0x800)
        int prim = 0; int64_t best = 0;
        for (int i = 0; i < nseg; i++) if (ps[i].q1 - ps[i].q0 > best) { best = ps[i].q1 - ps[i].q0; prim = i; }
        double rate = profile == 1 ? 0.05 : 0.001;
        std::vector<uint64_t> rec_idx(nseg);
        uint8_t mapq_all; uint16_t nf = noise_flags(&mapq_all); nf &= (uint16_t)~0x10;
        for (int i = 0; i < nseg; i++) {
            Piece& p = ps[i];
            // clips in alignment orientation: on the reverse strand left/right swap
            int64_t lq = p.strand > 0 ? p.q0 : len - p.q1, rq = p.strand > 0 ? len - p.q1 : p.q0;
            uint32_t clip_op = i == prim ? OP_S : OP_H;
            p.cb = cigar.size();
            if (lq) cigar.push_back(cg((uint32_t)lq, clip_op));
            p.refspan = body(p.q1 - p.q0, rate, 0, nullptr, 0);
            if (rq) cigar.push_back(cg((uint32_t)rq, clip_op));
            p.ce = cigar.size(); p.nm = (uint32_t)rng.below(200);
            uint16_t flag = (uint16_t)(nf | (p.strand < 0 ? 0x10 : 0) | (i == prim ? 0 : 0x800));
            rec_idx[i] = recs.size();
            push_rec(p.tid, (int32_t)p.pos, flag, mapq_all, p.cb, qid);
        }
        // SA strings: for record i list the others, primary first then by query order (samtools convention)
        for (int i = 0; i < nseg; i++) {
            Rec& r = recs[rec_idx[i]]; r.sa_kind = 1; r.sa_begin = sa.size();
            std::vector<int> order; if (i != prim) order.push_back(prim);
            for (int j = 0; j < nseg; j++) if (j != i && j != prim) order.push_back(j);
            char buf[256];
            for (int j : order) {
                const Piece& p = ps[j];
                int64_t lq = p.strand > 0 ? p.q0 : len - p.q1, rq = p.strand > 0 ? len - p.q1 : p.q0;
                int64_t m = p.q1 - p.q0, d = p.refspan - m; std::string cs;
                if (lq) cs += std::to_string(lq) + "S";
                cs += std::to_string(d < 0 ? m + d : m) + "M";
                if (d > 0) cs += std::to_string(d) + "D"; else if (d < 0) cs += std::to_string(-d) + "I";
                if (rq) cs += std::to_string(rq) + "S";
                int n = snprintf(buf, sizeof buf, "%s,%lld,%c,%s,%d,%u;", kNames[p.tid], (long long)p.pos + 1, p.strand > 0 ? '+' : '-', cs.c_str(), 60, p.nm);
                sa.insert(sa.end(), buf, buf + n);
            }
            r.sa_end = sa.size();
        }
    }
};

}  // namespace

extern "C" {

struct exlr_synth_out {
    uint64_t n_reads, n_ops, n_sa_bytes;
    uint32_t* cigar; uint64_t* cigar_off; int32_t* pos; int32_t* tid; uint16_t* flag; uint8_t* mapq; uint8_t* sa_kind;
    uint32_t* sa_off; uint8_t* sa_bytes; uint64_t* qid;
};

int exlr_synth_n_ref(void) { return 25; }
const char* exlr_synth_ref_name(int i) { return (i >= 0 && i < 25) ? kNames[i] : ""; }
int64_t exlr_synth_ref_len(int i) { return (i >= 0 && i < 25) ? kLens[i] : 0; }

// profile: 0 HiFi, 1 ONT, 2 split-heavy.  n_molecules: molecules (records >= molecules).  n_ultra: ONT reads forced to ~1 Mb.
int exlr_synth_generate(int profile, uint64_t seed, uint64_t n_molecules, int chr20_only, uint64_t n_ultra, exlr_synth_out* out)
{
    if (!out || profile < 0 || profile > 2) return -1;
    Gen g(profile, seed, chr20_only != 0);
    if (profile != 2) g.make_sites(profile == 0 ? 20 : 40);
    g.recs.reserve((size_t)(n_molecules * (profile == 2 ? 4 : 1.1)));
    uint64_t large_ins_every = n_molecules >= 1000 ? n_molecules / 10 : 0;      // ~10 large-INS cases per run
    const int64_t clip_targets[4] = {999, 1000, 1001, 5000};
    for (uint64_t m = 0; m < n_molecules; m++) {
        if (profile == 2) {
            int nseg = 2; while (nseg < 9 && g.rng.uni() < 0.5) nseg++;
            g.split(m, nseg, 0, 0);
        } else if (large_ins_every && m % large_ins_every == large_ins_every / 2) {
            g.split(m, 2, 1, clip_targets[(m / large_ins_every) % 4]);
        } else if (m < n_ultra) {
            g.linear(m, 1000000 - (int64_t)g.rng.below(50000));
        } else if (g.rng.uni() < g.sa_frac * 0.4) {        // 5 % of *records* carry SA: molecules average 2.5 records
            int nseg = 2 + (int)g.rng.below(3);
            g.split(m, nseg, 0, 0);
        } else {
            g.linear(m, 0);
        }
    }
    // coordinate sort (unmapped tid -1 last), stable so a molecule's records keep their order on ties
    std::vector<uint32_t> perm(g.recs.size());
    for (size_t i = 0; i < perm.size(); i++) perm[i] = (uint32_t)i;
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) {
        const Rec& x = g.recs[a]; const Rec& y = g.recs[b];
        uint32_t tx = (uint32_t)x.tid, ty = (uint32_t)y.tid;
        if (tx != ty) return tx < ty;
        return x.pos < y.pos; });
    uint64_t R = g.recs.size(), C = g.cigar.size(), A = g.sa.size();
    out->n_reads = R; out->n_ops = C; out->n_sa_bytes = A;
    out->cigar = (uint32_t*)malloc((C + 4) * 4); out->cigar_off = (uint64_t*)malloc((R + 1) * 8);
    out->pos = (int32_t*)malloc((R + 1) * 4); out->tid = (int32_t*)malloc((R + 1) * 4);
    out->flag = (uint16_t*)malloc((R + 1) * 2); out->mapq = (uint8_t*)malloc(R + 1); out->sa_kind = (uint8_t*)malloc(R + 1);
    out->sa_off = (uint32_t*)malloc((R + 1) * 4); out->sa_bytes = (uint8_t*)malloc(A + 16); out->qid = (uint64_t*)malloc((R + 1) * 8);
    if (!out->cigar || !out->cigar_off || !out->pos || !out->tid || !out->flag || !out->mapq || !out->sa_kind || !out->sa_off || !out->sa_bytes || !out->qid) return -3;
    if (A >= 0x7fffffffull) return -4;
    uint64_t co = 0; uint32_t so = 0;
    for (uint64_t i = 0; i < R; i++) {
        const Rec& r = g.recs[perm[i]];
        out->cigar_off[i] = co; out->sa_off[i] = so;
        uint64_t n = r.cig_end - r.cig_begin;
        if (n) memcpy(out->cigar + co, g.cigar.data() + r.cig_begin, n * 4);
        co += n;
        uint64_t s = r.sa_end - r.sa_begin;
        if (s) memcpy(out->sa_bytes + so, g.sa.data() + r.sa_begin, s);
        so += (uint32_t)s;
        out->pos[i] = r.pos; out->tid[i] = r.tid; out->flag[i] = r.flag; out->mapq[i] = r.mapq; out->sa_kind[i] = r.sa_kind; out->qid[i] = r.qid;
    }
    out->cigar_off[R] = co; out->sa_off[R] = so;
    return 0;
}

void exlr_synth_free(exlr_synth_out* o)
{
    if (!o) return;
    free(o->cigar); free(o->cigar_off); free(o->pos); free(o->tid); free(o->flag); free(o->mapq); free(o->sa_kind);
    free(o->sa_off); free(o->sa_bytes); free(o->qid);
    memset(o, 0, sizeof(*o));
}

}  // extern "C"
