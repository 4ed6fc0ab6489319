// exlr_kernels.cu — hand-written sm_100a kernels for excord-lr's signal-extraction hot path.
//
//   kernel 0  k0_classify    record filter (main.rs:169-190) + tid check (:198) + ordered list of
//                            kept records carrying an SA aux (:206), one chained scan
//   kernel 1a k1a_screen     streams the CIGAR array once at HBM speed and lists the 512-op steps that
//                            hold an I/D >= indel_min (main.rs:553,569) or an unknown op code
//   kernel 1b k1b_claim      listed steps -> records (one owner per record), short / long lists
//             k1b_walk       one thread per short record: the CIGAR scan of main.rs:523-600 and the
//                            merge predicates of :612-635, :673-678, all thread-local
//   kernel 1c k1_flat(list)  the long records, one per tile, by the flat block scan below
//   kernel 1  k1_flat        CIGAR scan of everything (event-dense batches, EXLR_OPT_CIGAR_KERNEL=2): the
//                            CIGAR stream is staged through shared memory by TMA bulk copies
//                            (cp.async.bulk + mbarrier, 5-stage ring) and scanned flat — one block prefix
//                            sum of the reference-consuming lengths, record boundaries resolved from the
//                            staged offsets — so load balance does not depend on CIGAR lengths.
//             k1_warp        warp-per-record variant (kept for A/B measurement)
//   kernel 3a k3a_sa_cigar   per-op-type sums + first-match offset of SA records' own CIGARs
//                            (main.rs:214-306, utils.rs:12-42), 2/4/8-lane group per record
//   kernel 3b k3b_sa_events  SA parse (utils.rs:88-139), -k cap (main.rs:311), stable segment sort
//                            (main.rs:322), large-INS rules (:340-486), split pairs (:488-516)
//   kernel 4a k4a_line_scan  pair-merge rule (main.rs:612-635) + far-edge domain check (:673-678)
//                            folded into the per-record line count, chained scan -> line offsets
//   kernel 4b k4b_place      ordered compaction: every event lands at its reference output
//                            position as a 48-byte exlr_event
//   kernel 5a k5a_line_bytes byte length of every output line -> chained scan -> byte offsets  (optional)
//   kernel 5b k5b_format     the lines themselves (utils.rs:225-236, 269-280)                  (optional)
//
// Consecutive kernels of a chain use programmatic dependent launch (griddep_wait / griddep_launch below).
// HBM-bound integer/byte work: no tensor cores anywhere on this path.
#include <cstdint>
#include <cuda_runtime.h>

#include "exlr_device.cuh"

namespace exlr {

// ======================================================================================
// small helpers
// ======================================================================================
__device__ __forceinline__ bool keep_record(const DevParams& P, uint32_t flag, uint32_t mapq)
{
    if (P.exclude_secondary && (flag & 0x100u)) return false;   // main.rs:169
    if (P.exclude_unmapped && (flag & 0x4u)) return false;      // main.rs:174
    if (mapq < P.mapq) return false;                            // main.rs:179
    if (flag & P.exclude_flag) return false;                    // main.rs:185
    return true;
}

// The word holds ~key so that a zeroing memset means "no error" and atomicMax keeps the smallest key.
__device__ __forceinline__ void report(Ctrl* c, uint32_t read, uint32_t rank)
{
    atomicMax(&c->err_key, ~(((unsigned long long)read << 8) | rank));
}

__device__ __forceinline__ uint32_t abs_diff(uint32_t a, uint32_t b) { return a > b ? a - b : b - a; }
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may be placed on the
// SMs while its predecessor in the stream is still draining; griddep_wait() returns once the predecessor has completed and its
// writes are visible, so every global access of such a kernel comes after it.  griddep_launch() lets the successor be placed.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// ops consuming the reference in the indel arm: M(0) D(2) N(3) =(7)   (main.rs:528-545, 586-598)
__device__ __forceinline__ uint32_t consumes_ref(uint32_t op) { return (0x8Du >> op) & 1u; }

__device__ __forceinline__ void store_event(exlr_event* dst, int64_t ls, int64_t le, int64_t rs, int64_t re,
                                            uint32_t read, uint32_t lc, uint32_t rc, uint32_t meta)
{
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4((uint32_t)ls, (uint32_t)((uint64_t)ls >> 32), (uint32_t)le, (uint32_t)((uint64_t)le >> 32));
    d[1] = make_uint4((uint32_t)rs, (uint32_t)((uint64_t)rs >> 32), (uint32_t)re, (uint32_t)((uint64_t)re >> 32));
    d[2] = make_uint4(read, lc, rc, meta);
}

// EXLR_OPT_TRACE for the small kernels: {CTA start, a mid point, CTA end} by thread 0, entry blockIdx.x (mod 8192).
// EXLR_OPT_TRACE = 7 is the timeline mode: every kernel folds its CTAs' start / end into entry `id` with atomicMin / atomicMax
// ({~first CTA start, last CTA end, -, CTAs}; the buffer is zeroed per submit), which shows the gaps and overlaps of a whole step
// (tools/timeline.py).
struct CtaTrace {
    unsigned long long* d; unsigned long long t0, t1; bool tl;
    __device__ __forceinline__ CtaTrace(const DevBatch& B, uint32_t sel) : d(nullptr), t0(0), t1(0), tl(false)
    {
        if (!B.dbg) return;
        if (B.dbg_sel == 7u) { d = B.dbg + 4ull * sel; tl = true; }
        else if (B.dbg_sel == sel) d = B.dbg + 4ull * (blockIdx.x & 8191u);
        if (d) t0 = gtimer();
    }
    __device__ __forceinline__ void mid() { if (d && !t1) t1 = gtimer(); }
    __device__ __forceinline__ void end()
    {
        if (!d || threadIdx.x != 0) return;
        const unsigned long long t2 = gtimer();
        if (tl) { atomicMax(d, ~t0); atomicMax(d + 1, t2); atomicAdd(d + 3, 1ull); }
        else { d[0] = t0; d[1] = t1; d[2] = t2; d[3] = 0; }
    }
};

// ---- chained scan (decoupled look-back), one status word per tile: flag<<62 | value ----
// 256 threads x 16 records = 4096 records per tile; the look-back is done by a whole warp, 32 tiles at a time.
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 16;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
static constexpr unsigned long long ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, ST_VALUE = (1ull << 62) - 1;

// All SCAN_THREADS threads: publish this tile's aggregate, return the sum of all previous tiles.
// The look-back window is the whole CTA (256 predecessors per probe, every status word fetched by its own thread), so a
// batch of a few hundred tiles resolves in one L2 round trip instead of a serial walk of 32-tile windows.
// Tile ids come from a ticket, so every predecessor is running or done and publishes without waiting for anybody.
__device__ __forceinline__ uint32_t chained_prefix(unsigned long long* status, uint32_t tile, uint32_t agg,
                                                   uint32_t* s_sum /*[8]*/, uint32_t* s_has /*[8]*/)
{
    volatile unsigned long long* st = status;
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (tile == 0) { if (t == 0) st[0] = ST_PREFIX | agg; return 0; }
    if (t == 0) st[tile] = ST_AGG | agg;
    uint32_t prefix = 0;
    for (int base = (int)tile - 1;; base -= SCAN_THREADS) {
        const int idx = base - (int)t;
        unsigned long long s = ST_PREFIX;                                      // tiles before 0: prefix 0
        if (idx >= 0) { do { s = st[idx]; } while ((s >> 62) == 0ull); }
        const uint32_t pm = __ballot_sync(0xffffffffu, (uint32_t)(s >> 62) == 2u);
        const uint32_t fp = pm ? (uint32_t)(__ffs(pm) - 1) : 31u;             // nearest inclusive prefix inside this warp's window
        uint32_t v = lane <= fp ? (uint32_t)(s & ST_VALUE) : 0u;
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) { s_sum[w] = v; s_has[w] = pm != 0u; }
        __syncthreads();
        bool found = false;
#pragma unroll
        for (int k = 0; k < SCAN_THREADS / 32; k++) { if (!found) { prefix += s_sum[k]; found = s_has[k] != 0u; } }
        __syncthreads();                                                       // s_sum / s_has are reused by the next window
        if (found) break;
    }
    if (t == 0) st[tile] = ST_PREFIX | (unsigned long long)(agg + prefix);
    return prefix;
}

// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp /*[8]*/, uint32_t* total)
{
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < SCAN_THREADS / 32; i++) { uint32_t x = s_warp[i]; if ((uint32_t)i < w) base += x; tot += x; }
    *total = tot;
    return base + incl - v;
}

// tile prefix: block scan + chained look-back; returns this thread's exclusive prefix over the whole batch
__device__ __forceinline__ uint32_t tile_excl_scan(unsigned long long* status, uint32_t tile, uint32_t mine,
                                                   uint32_t* s_warp /*[16]*/, uint32_t* grand_total_if_last)
{
    uint32_t total;
    const uint32_t excl = block_excl_scan(mine, s_warp, &total);
    __syncthreads();                                                           // s_warp is reused by the look-back
    const uint32_t p = chained_prefix(status, tile, total, s_warp, s_warp + 8);
    *grand_total_if_last = p + total;
    return p + excl;
}

// ======================================================================================
// kernel 0: filter + tid check + ordered SA-record list
// ======================================================================================
__global__ void __launch_bounds__(SCAN_THREADS) k0_classify(DevBatch B, DevParams P)
{
    __shared__ uint32_t s_tile, s_warp[16];
    griddep_launch();                                  // kernel 3a may be placed; it waits for this grid before it reads anything
    CtaTrace tr(B, 2);
    if (threadIdx.x == 0) s_tile = atomicAdd(&B.ctrl->ticket_a, 1u);
    __syncthreads();
    const uint32_t tile = s_tile, n = B.n_reads;
    const uint32_t r0 = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t sa_mask = 0, kept = 0;
    if (r0 + SCAN_ITEMS <= n) {
        // 16 consecutive records per thread: 128-bit loads of every per-record array
        union { uint4 v[2]; uint16_t h[16]; } fl;
        union { uint4 v; uint8_t b[16]; } mq, kd;
        union { uint4 v[4]; int32_t i[16]; } td;
        fl.v[0] = __ldg(reinterpret_cast<const uint4*>(B.flag + r0)); fl.v[1] = __ldg(reinterpret_cast<const uint4*>(B.flag + r0) + 1);
        mq.v = __ldg(reinterpret_cast<const uint4*>(B.mapq + r0));
        kd.v = __ldg(reinterpret_cast<const uint4*>(B.sa_kind + r0));
#pragma unroll
        for (int k = 0; k < 4; k++) td.v[k] = __ldg(reinterpret_cast<const uint4*>(B.tid + r0) + k);
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            if (keep_record(P, fl.h[i], mq.b[i])) {
                kept++;
                const int32_t t = td.i[i];
                if (t < 0 || t >= B.n_ref) report(B.ctrl, r0 + i, RANK_TID);     // record.contig() panics (main.rs:198)
                else if (kd.b[i] != EXLR_SA_NONE) sa_mask |= 1u << i;
            }
        }
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int k = 0; k < 4; k++) reinterpret_cast<uint4*>(B.csa + r0)[k] = z;
        if (P.split_only) {                                                      // kernel 1 does not run (main.rs:523)
#pragma unroll
            for (int k = 0; k < 8; k++) reinterpret_cast<uint4*>(B.k1 + r0)[k] = z;
        }
    } else {
        for (int i = 0; i < SCAN_ITEMS; i++) {
            const uint32_t r = r0 + i;
            if (r >= n) break;
            if (keep_record(P, B.flag[r], B.mapq[r])) {
                kept++;
                const int32_t t = B.tid[r];
                if (t < 0 || t >= B.n_ref) report(B.ctrl, r, RANK_TID);
                else if (B.sa_kind[r] != EXLR_SA_NONE) sa_mask |= 1u << i;
            }
            B.csa[r] = 0;
            if (P.split_only) B.k1[r] = make_uint2(0u, 0u);
        }
    }
    uint32_t grand;
    tr.mid();
    uint32_t at = tile_excl_scan(B.scan_a, tile, (uint32_t)__popc(sa_mask), s_warp, &grand);
    while (sa_mask) { const int i = __ffs(sa_mask) - 1; sa_mask &= sa_mask - 1; B.sa_list[at++] = r0 + i; }
    // kept counter: one atomic per warp
    for (int d = 16; d; d >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, d);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(&B.ctrl->n_kept, kept);
    if (threadIdx.x == 0 && tile == (n + SCAN_TILE - 1) / SCAN_TILE - 1) B.ctrl->n_sa = grand;
    tr.end();
}


// ======================================================================================
// kernel 1 (variant B): warp per record
// ======================================================================================
// One record scanned by the 32 lanes of a warp, 32 ops per step (all lanes of the warp must call).
__device__ __forceinline__ void k1_warp_record(const DevBatch& B, const DevParams& P, uint32_t r)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t flag = B.flag[r], mq = B.mapq[r];
    if (!keep_record(P, flag, mq)) { if (lane == 0) B.k1[r] = make_uint2(0u, 0u); return; }
    const unsigned long long o0 = B.cigar_off[r], o1 = B.cigar_off[r + 1];
    const uint32_t pos2 = (uint32_t)B.pos[r];
    uint32_t carry = 0, cnt = 0, info = 0;
    uint32_t pL = 0, pn = 0, pdel = 0;          // previous event of this record (warp-uniform)
    for (unsigned long long b = o0; b < o1; b += 32) {
        const bool valid = b + lane < o1;
        const uint32_t v = valid ? __ldg(B.cigar + b + lane) : 0u;
        const uint32_t op = v & 15u, len = v >> 4;
        if (valid && op > 8u) report(B.ctrl, r, RANK_CIGAR_OP);
        const uint32_t c = (valid && op <= 8u && consumes_ref(op)) ? len : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
        const uint32_t L = carry + incl - c;
        const bool isev = valid && (op == 1u || op == 2u) && len >= P.indel_min;
        const uint32_t bal = __ballot_sync(0xffffffffu, isev);
        if (bal) {
            const uint32_t below = bal & ((1u << lane) - 1u);
            const uint32_t rank = __popc(below);
            const int src = below ? 31 - __clz(below) : 0;
            uint32_t qL = __shfl_sync(0xffffffffu, L, src), qn = __shfl_sync(0xffffffffu, len, src),
                     qdel = __shfl_sync(0xffffffffu, (uint32_t)(op == 2u), src);
            bool has_prev = below != 0;
            if (!has_prev && cnt) { qL = pL; qn = pn; qdel = pdel; has_prev = true; }
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&B.ctrl->n_raw, (uint32_t)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            uint32_t myflags = 0;
            if (isev) {
                const uint32_t seq = cnt + rank, del = op == 2u;
                if (has_prev && del && qdel) {
                    if (seq == 1 && abs_diff(pos2 + L, pos2 + qL + qn) < P.merge_min) myflags |= K1_PAIR_MERGE;   // main.rs:615
                    if (abs_diff(pos2 + qL, pos2 + L + len) < P.merge_min) myflags |= K1_FAR_HIT;                 // main.rs:673-678
                }
                const uint32_t slot = B.prim_slots + base + rank;
                if (slot < B.raw_cap) {
                    uint4* d = reinterpret_cast<uint4*>(B.raw + slot);
                    d[0] = make_uint4(r, seq, L, len | (del << 31));
                    d[1] = make_uint4(has_prev ? qL : 0u, 0u, 0u, 0u);
                } else B.ctrl->overflow = 1;
            }
            for (int d = 16; d; d >>= 1) myflags |= __shfl_xor_sync(0xffffffffu, myflags, d);
            info |= myflags;
            const int last = 31 - __clz(bal);
            pL = __shfl_sync(0xffffffffu, L, last); pn = __shfl_sync(0xffffffffu, len, last);
            pdel = __shfl_sync(0xffffffffu, (uint32_t)(op == 2u), last);
            cnt += __popc(bal);
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) B.k1[r] = make_uint2(carry, (cnt & K1_CNT_MASK) | info);
}

__global__ void __launch_bounds__(256) k1_warp(DevBatch B, DevParams P)
{
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < B.n_reads; r += warps) k1_warp_record(B, P, r);
}

// ======================================================================================
// kernel 1 (default): flat TMA-staged block scan over the CIGAR stream
// ======================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "EXLR_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra EXLR_DONE;\n"
                 "bra EXLR_WAIT;\n"
                 "EXLR_DONE:\n"
                 "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

static constexpr int K1_THREADS = 128;
static constexpr int K1_WARPS = K1_THREADS / 32;
static constexpr int K1_V = 16;                        // consecutive ops per thread per scan step
static constexpr int K1_CHUNK = K1_THREADS * K1_V;     // ops per TMA bulk copy = per scan step (8 KB)
static constexpr int K1_STAGES = 5;                    // ring of bulk-copy stages (40 KB), kept full across tiles
static constexpr int K1_MAX_RPC = K1_THREADS;          // records per tile (thread t owns record t)
static constexpr int K1_CAP = K1_THREADS;              // staged events per flush (one per thread)
static constexpr int K1_MAX_TILES = 256;               // tiles per CTA
static constexpr uint32_t K1_SCREEN_CHUNKS = 2;         // tiles resident in at most this many stages are screened for events first

struct __align__(16) K1Stage { uint32_t fp, pexcl, n_type, pad; };

struct __align__(128) K1Smem {
    uint32_t buf[K1_STAGES][K1_CHUNK];                 // CIGAR ops as landed by TMA; overwritten in place by per-op prefixes
    K1Stage stage[K1_CAP];
    unsigned long long tb[K1_MAX_TILES][2];            // [first, last) op offset of every tile of this CTA
    uint32_t tidx[K1_MAX_TILES];                       // tile ids of this CTA: b, b + grid, ... or its share of the long-record list
    uint32_t rpos[K1_MAX_RPC];                         // record.pos() as u32 (aligments_event.rs:43)
    uint32_t roff[K1_MAX_RPC + 1];
    uint32_t pstart[K1_MAX_RPC + 1];
    uint32_t rcnt[K1_MAX_RPC];
    uint32_t rflags[K1_MAX_RPC];
    uint32_t fhead[K1_MAX_RPC];
    uint32_t rkeep[K1_MAX_RPC];
    uint32_t tpre[2][K1_THREADS];                      // warp-relative exclusive prefix of each thread's 16 ops
    uint32_t wtot[2][K1_WARPS];
    uint32_t evcnt[K1_WARPS];
    K1Stage carry_ev;
    uint32_t has_carry, gbase, flushes, in_slab;
    unsigned long long full[K1_STAGES];
};

// largest i in [0, nr) with roff[i] <= fp (records are back to back; empty records share a start and the
// last of them owns the ops)
__device__ __forceinline__ uint32_t k1_find_read(const uint32_t* roff, uint32_t nr, uint32_t fp)
{
    uint32_t lo = 0, hi = nr;               // invariant: roff[lo] <= fp, answer in [lo, hi)
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (roff[mid] <= fp) lo = mid; else hi = mid; }
    return lo;
}

// Where a tile's raw events go: the tile's first flush lands in its own fixed slice of the raw buffer
// (no atomic, no round trip); anything beyond goes to the shared overflow region behind the slices.
struct K1Out { RawEv* raw; Ctrl* ctrl; uint32_t* tile_cnt; uint32_t raw_cap, merge_min, prim_slots, capt_log2, slab; };

// Resolve and write out the `m` staged events (m <= K1_CAP, one per thread).  Block-uniform call.
// `spare` (meaningful in thread 0 only) is an overflow slab of K1_CAP slots reserved ahead of time, so that no flush
// ever waits for an atomic: the flush that consumes it immediately reserves the next one.
__device__ __forceinline__ void k1_flush(K1Smem& S, const K1Out& O, uint32_t tile, uint32_t ra, uint32_t nr, uint32_t m, uint32_t& spare)
{
    const uint32_t t = threadIdx.x;
    if (t == 0) {
        const uint32_t capt = O.prim_slots ? (1u << O.capt_log2) : 0u;
        if (S.flushes == 0 && m <= capt) { S.gbase = tile << O.capt_log2; O.tile_cnt[tile] = m; S.in_slab = 0; }
        else {
            if (S.flushes == 0) O.tile_cnt[tile] = 0;                     // slice unused: everything of this tile overflows
            if (m <= O.slab) {
                S.gbase = O.prim_slots + spare; S.in_slab = 1;
                spare = atomicAdd(&O.ctrl->n_raw, O.slab);                // not needed before the next overflowing flush
            } else {                                                      // larger than a slab: exact, synchronous reservation
                S.gbase = O.prim_slots + atomicAdd(&O.ctrl->n_raw, m); S.in_slab = 0;
            }
        }
        S.flushes++;
    }
    __syncthreads();                                    // staging, pstart, gbase visible
    K1Stage ev; uint32_t i = 0; bool has_prev = false; K1Stage pv;
    pv.fp = 0; pv.pexcl = 0; pv.n_type = 0; pv.pad = 0; ev = pv;
    if (t < m) {
        ev = S.stage[t];
        i = k1_find_read(S.roff, nr, ev.fp);
        if (t > 0) { pv = S.stage[t - 1]; has_prev = pv.fp >= S.roff[i]; }
        else if (S.has_carry) { pv = S.carry_ev; has_prev = pv.fp >= S.roff[i]; }
        if (t == 0 || S.stage[t - 1].fp < S.roff[i]) S.fhead[i] = t;      // first event of record i in this flush
    }
    __syncthreads();
    uint32_t seq = 0;
    const uint32_t slot = S.gbase + t;
    if (t < m) {
        seq = S.rcnt[i] + t - S.fhead[i];
        const uint32_t L = ev.pexcl - S.pstart[i];
        const uint32_t len = ev.n_type & 0x7fffffffu, del = ev.n_type >> 31;
        uint32_t prevL = 0;
        if (has_prev) {
            prevL = pv.pexcl - S.pstart[i];
            const uint32_t pn = pv.n_type & 0x7fffffffu, pdel = pv.n_type >> 31;
            if (del && pdel) {
                const uint32_t pos2 = S.rpos[i];
                uint32_t fl = 0;
                if (seq == 1 && abs_diff(pos2 + L, pos2 + prevL + pn) < O.merge_min) fl |= K1_PAIR_MERGE;   // main.rs:615
                if (abs_diff(pos2 + prevL, pos2 + L + len) < O.merge_min) fl |= K1_FAR_HIT;                 // main.rs:673-678
                if (fl) atomicOr(&S.rflags[i], fl);
            }
        }
        if (slot < O.raw_cap) {
            uint4* d = reinterpret_cast<uint4*>(O.raw + slot);
            d[0] = make_uint4(S.rkeep[i] ? ra + i : 0xffffffffu, seq, L, ev.n_type);
            d[1] = make_uint4(prevL, 0u, 0u, 0u);
        } else O.ctrl->overflow = 1;
    } else if (S.in_slab && t < O.slab && slot < O.raw_cap) {
        reinterpret_cast<uint4*>(O.raw + slot)[0] = make_uint4(0xffffffffu, 0u, 0u, 0u);   // unused slab slot
    }
    __syncthreads();                                    // every seq computed before rcnt moves
    if (t < m) {
        const bool tail = (t == m - 1) || (S.stage[t + 1].fp >= S.roff[i + 1]);
        if (tail) S.rcnt[i] = seq + 1;
        if (t == m - 1) { S.carry_ev = ev; S.has_carry = 1; }
    }
    __syncthreads();
}

// Rare path of one scan step: rank this step's events over the CTA and stage them (flushing as needed).
// Block-uniform call.  evm: bit k set = my k-th op is an event; the per-op prefixes are read back from `pre`.
__device__ __forceinline__ uint32_t k1_stage_events(K1Smem& S, const K1Out& O, uint32_t tile, uint32_t ra, uint32_t nr,
                                                    uint32_t staged, uint32_t evm, const uint32_t* v, const uint32_t* pre,
                                                    uint32_t fp0, uint32_t pbase, uint32_t& spare)
{
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
    const uint32_t nev = __popc(evm);
    uint32_t evincl = nev;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, evincl, d); if (lane >= (uint32_t)d) evincl += o; }
    if (lane == 31) S.evcnt[w] = evincl;
    __syncthreads();
    uint32_t evbase = 0, evtotal = 0;
#pragma unroll
    for (int k = 0; k < K1_WARPS; k++) { const uint32_t x = S.evcnt[k]; if ((uint32_t)k < w) evbase += x; evtotal += x; }
    if (staged + evtotal > K1_CAP && staged) { k1_flush(S, O, tile, ra, nr, staged, spare); staged = 0; }
    const uint32_t my0 = evbase + evincl - nev;                      // rank of my first event in this step
    if (evtotal <= K1_CAP) {                                          // the usual case: everything fits the staging area
        if (evm) {
            uint32_t at = staged + my0;
#pragma unroll
            for (int k = 0; k < K1_V; k++) {                             // static indices keep v[] in registers
                if (evm & (1u << k)) {
                    K1Stage x;
                    x.fp = fp0 + k; x.pexcl = pbase + pre[k];
                    x.n_type = (v[k] >> 4) | (((v[k] & 15u) == 2u) ? 0x80000000u : 0u);
                    x.pad = 0;
                    S.stage[at++] = x;
                }
            }
        }
        return staged + evtotal;
    }
    for (uint32_t round0 = 0; round0 < evtotal; round0 += K1_CAP) {  // event-dense step: K1_CAP events at a time
        uint32_t rk = my0;
#pragma unroll
        for (int k = 0; k < K1_V; k++) {
            if (!(evm & (1u << k))) continue;
            if (rk >= round0 && rk < round0 + K1_CAP) {
                K1Stage x;
                x.fp = fp0 + k; x.pexcl = pbase + pre[k];
                x.n_type = (v[k] >> 4) | (((v[k] & 15u) == 2u) ? 0x80000000u : 0u);
                x.pad = 0;
                S.stage[staged + rk - round0] = x;
            }
            rk++;
        }
        k1_flush(S, O, tile, ra, nr, staged + min((uint32_t)K1_CAP, evtotal - round0), spare);
        staged = 0;
    }
    return staged;
}

// op code -> flag byte, looked up with one PRMT (selector nibble 0 = op code):
//   bit0 = consumes the reference in the indel arm (M D N =, main.rs:528-545)   bit1 = I or D (event candidate)
// Bytes 1..7 carry bit7 so that selectors 9..15 (PRMT sign-replicate mode) read 0xff: bit6 then flags an unknown
// op code; selector 8 (X) replicates the clear msb of byte 0 and reads 0x00: X neither consumes nor is an event.
static constexpr uint32_t K1_LUT_LO = 0x81838201u;     // N D I M
static constexpr uint32_t K1_LUT_HI = 0x81808080u;     // = P H S

// Persistent: CTA b scans tiles b, b + gridDim, b + 2 gridDim, ... (`rpc` records each; interleaved so that a
// run of event-dense tiles is spread over many CTAs).  All its tile boundaries are fetched up front, so thread 0
// keeps the TMA ring K1_STAGES chunks ahead of the scan ACROSS tile boundaries; the next tile's per-record
// offsets, positions and filter inputs are prefetched into registers while the current tile is scanned.
__global__ void __launch_bounds__(K1_THREADS, 4) k1_flat(DevBatch B, DevParams P, uint32_t rpc, uint32_t n_tiles, const uint32_t* list)
{
    extern __shared__ __align__(128) unsigned char k1_smem_raw[];
    K1Smem& S = *reinterpret_cast<K1Smem*>(k1_smem_raw);
    const uint32_t t = threadIdx.x, lane = t & 31, w = t >> 5;
    // listed mode (kernel 1c): the tiles are the long records kernel 1b put on its list, one record each (rpc = 1);
    // every one of them holds an event candidate, so nothing is screened
    if (list) { griddep_wait(); n_tiles = B.ctrl->n_long; }
    if (blockIdx.x >= n_tiles) return;
    CtaTrace tr(B, 11);
    const bool trace = B.dbg && B.dbg_sel == 1u && !list;
    const unsigned long long tr_start = trace ? gtimer() : 0ull; unsigned long long tr_first = 0; uint32_t tr_scanned = 0;
    const uint32_t ntile = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;         // <= K1_MAX_TILES (host)
    for (uint32_t k = t; k < ntile; k += K1_THREADS) {
        const uint32_t id = blockIdx.x + k * gridDim.x;
        const uint32_t tile = list ? list[id] : id;
        S.tidx[k] = tile;
        const unsigned long long r0 = (unsigned long long)tile * rpc;
        S.tb[k][0] = B.cigar_off[min(r0, (unsigned long long)B.n_reads)];
        S.tb[k][1] = B.cigar_off[min(r0 + rpc, (unsigned long long)B.n_reads)];
    }
    uint32_t spare = 0;                                    // thread 0: reserved overflow slab (see k1_flush)
    if (t == 0) {
        for (int s = 0; s < K1_STAGES; s++) mbar_init(&S.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        spare = atomicAdd(&B.ctrl->n_raw, B.slab);
    }
    // per-record inputs of the next tile that is known to need the full scan, prefetched into registers (see the tile loop)
    unsigned long long pre_off = 0; uint32_t pre_flag = 0, pre_mapq = 0, pre_pos = 0, pre_tile = 0xffffffffu;
    __syncthreads();

    // issue cursor (thread 0): next chunk to request = chunk ic of tile it_i; gi = chunks requested so far
    uint32_t it_i = 0, ic = 0, gi = 0;
    auto issue_upto = [&](uint32_t limit) {                // request chunks while fewer than `limit` have been requested
        while (gi < limit && it_i < ntile) {
            const unsigned long long oa4 = S.tb[it_i][0] & ~3ull;
            const uint32_t span_hi = (uint32_t)(S.tb[it_i][1] - oa4);
            const uint32_t nch = (span_hi + K1_CHUNK - 1) / K1_CHUNK;
            if (ic >= nch) { it_i++; ic = 0; continue; }
            const uint32_t first = ic * K1_CHUNK;
            const uint32_t nops = min((uint32_t)K1_CHUNK, span_hi - first);
            const uint32_t bytes = ((nops * 4u) + 15u) & ~15u;          // the cigar buffer is padded by 16 bytes
            unsigned long long* bar = &S.full[gi % K1_STAGES];
            mbar_expect_tx(bar, bytes);
            tma_load_1d(S.buf[gi % K1_STAGES], B.cigar + oa4 + first, bytes, bar);
            gi++; ic++;
        }
    };
    if (t == 0) issue_upto(K1_STAGES);

    const uint32_t imin16 = P.indel_min >= (1u << 28) ? 0xffffffffu : (P.indel_min << 4);
    const K1Out O{B.raw, B.ctrl, B.tile_cnt, B.raw_cap, P.merge_min, B.prim_slots, B.capt_log2, B.slab};
    uint32_t gs = 0;                                       // scan steps done by this CTA (= chunks consumed)
    for (uint32_t it = 0; it < ntile; it++) {
        const uint32_t tile = S.tidx[it];
        const uint32_t ra = tile * rpc, nr = min(rpc, B.n_reads - ra);
        const unsigned long long oa = S.tb[it][0], ob = S.tb[it][1], oa4 = oa & ~3ull;
        const uint32_t span_lo = (uint32_t)(oa - oa4), span_hi = (uint32_t)(ob - oa4);
        if (ob - oa4 >= 0x80000000ull) { if (t == 0) B.ctrl->overflow = 1; return; }    // block-uniform
        const uint32_t nchunks = (span_hi + K1_CHUNK - 1) / K1_CHUNK;
        // ---- event-free tiles leave early ------------------------------------------------------------------------
        // total_consume and the prefixes are only ever read for records that have an indel event (kernel 4b), and events
        // (I/D >= indel_min) are sparse: a tile that is fully resident (<= 2 stages) is first screened with 4 instructions per op
        // -- table look-up, OR, compare, predicated OR.  A clean tile gets zeroed summaries and is done; anything suspicious
        // (event candidate, unknown op code) takes the full scan below, which re-reads the same, untouched stages.
        if (nchunks <= K1_SCREEN_CHUNKS && !list) {
            uint32_t sus = 0, flags = 0;
            for (uint32_t c = 0; c < nchunks; c++) {
                mbar_wait(&S.full[(gs + c) % K1_STAGES], ((gs + c) / K1_STAGES) & 1u);
                const uint4* m4 = reinterpret_cast<const uint4*>(S.buf[(gs + c) % K1_STAGES]) + t * (K1_V / 4);
                const uint32_t fp0 = c * K1_CHUNK + t * K1_V;
                const bool edge = fp0 < span_lo || fp0 + K1_V > span_hi;
                uint32_t vv[K1_V];
#pragma unroll
                for (int k = 0; k < K1_V / 4; k++) { const uint4 q = m4[k]; vv[4 * k] = q.x; vv[4 * k + 1] = q.y; vv[4 * k + 2] = q.z; vv[4 * k + 3] = q.w; }
                if (edge) {                                                  // first / last thread of the tile only
#pragma unroll
                    for (int k = 0; k < K1_V; k++) if (fp0 + k < span_lo || fp0 + k >= span_hi) vv[k] = 0u;
                }
#pragma unroll
                for (int k = 0; k < K1_V; k++) {
                    uint32_t f;
                    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(f) : "r"(K1_LUT_LO), "r"(K1_LUT_HI), "r"(vv[k]));
                    flags |= f;
                    asm("{\n.reg .pred p;\nsetp.ge.u32 p, %1, %2;\n@p or.b32 %0, %0, %3;\n}" : "+r"(sus) : "r"(vv[k]), "r"(imin16), "r"(f));
                }
            }
            if (trace && !tr_first) tr_first = gtimer();
            if (!__syncthreads_or((sus & 2u) | (flags & 0x40u))) {
                if (t < nr) B.k1[ra + t] = make_uint2(0u, 0u);
                if (t == 0) B.tile_cnt[tile] = 0;
                gs += nchunks;
                if (t == 0) issue_upto(gs + K1_STAGES);
                continue;
            }
        }
        tr_scanned++;
        // per-record state of the tile: from the register prefetch when this tile was known to need the scan (long records),
        // straight from global memory otherwise (a screened tile that turned out to hold events)
        if (pre_tile != it && t < nr) {
            pre_off = B.cigar_off[ra + t]; pre_flag = B.flag[ra + t]; pre_mapq = B.mapq[ra + t]; pre_pos = (uint32_t)B.pos[ra + t];
        }
        if (t < nr) {
            S.roff[t] = (uint32_t)(pre_off - oa4); S.rkeep[t] = keep_record(P, pre_flag, pre_mapq) ? 1u : 0u; S.rpos[t] = pre_pos;
            S.rcnt[t] = 0; S.rflags[t] = 0; S.fhead[t] = 0; S.pstart[t] = 0;
        }
        if (t == 0) { S.roff[nr] = span_hi; S.has_carry = 0; S.flushes = 0; }
        __syncthreads();
        if (it + 1 < ntile) {                              // next tile needs the scan for sure (long records, or listed): prefetch its inputs
            const unsigned long long na4 = S.tb[it + 1][0] & ~3ull;
            if (list || (uint32_t)((S.tb[it + 1][1] - na4 + K1_CHUNK - 1) / K1_CHUNK) > K1_SCREEN_CHUNKS) {
                const uint32_t ra2 = S.tidx[it + 1] * rpc, nr2 = min(rpc, B.n_reads - ra2);
                if (t < nr2) { pre_off = B.cigar_off[ra2 + t]; pre_flag = B.flag[ra2 + t]; pre_mapq = B.mapq[ra2 + t]; pre_pos = (uint32_t)B.pos[ra2 + t]; }
                pre_tile = it + 1;
            }
        }
        uint32_t carry = 0, staged = 0;
        for (uint32_t c = 0; c < nchunks; c++, gs++) {
            mbar_wait(&S.full[gs % K1_STAGES], (gs / K1_STAGES) & 1u);
            uint32_t* buf = S.buf[gs % K1_STAGES];
            uint4* mine4 = reinterpret_cast<uint4*>(buf) + t * (K1_V / 4);
            const uint32_t fp0 = c * K1_CHUNK + t * K1_V;                // flat position of my first op
            uint32_t v[K1_V];
#pragma unroll
            for (int k = 0; k < K1_V / 4; k++) {
                const uint4 q = mine4[k];
                v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
            }
            if (fp0 < span_lo || fp0 + K1_V > span_hi) {                 // tile edge: blank the ops outside my records
#pragma unroll
                for (int k = 0; k < K1_V; k++) if (fp0 + k < span_lo || fp0 + k >= span_hi) v[k] = 0u;
            }
            // decode; the thread-local exclusive prefixes replace the ops in shared memory as we go
            uint32_t tsum = 0, evm = 0, flags = 0;
#pragma unroll
            for (int g = 0; g < K1_V / 4; g++) {
                uint32_t e[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int k = 4 * g + q;
                    uint32_t f;                                          // prmt.b32, not __byte_perm: the intrinsic masks selector bit 3
                    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(f) : "r"(K1_LUT_LO), "r"(K1_LUT_HI), "r"(v[k]));
                    e[q] = tsum;
                    // tsum += consumes ? len : 0 as one multiply-add; evm |= (I or D) && len >= indel_min (main.rs:553,569)
                    // as one compare + one predicated LOP3.  Spelled in PTX so the 0/1 multiply is not turned into compare+select.
                    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(tsum) : "r"(f & 1u), "r"(v[k] >> 4));
                    const uint32_t bit = k ? (f << (k - 1)) : (f >> 1); // bit k <- "I or D"
                    asm("{\n.reg .pred p;\nsetp.ge.u32 p, %1, %2;\n@p lop3.b32 %0, %3, %4, %0, 0xEA;\n}"
                        : "+r"(evm) : "r"(v[k]), "r"(imin16), "r"(bit), "r"(1u << k));
                    flags |= f;
                }
                mine4[g] = make_uint4(e[0], e[1], e[2], e[3]);
            }
            uint32_t wincl = tsum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, wincl, d); if (lane >= (uint32_t)d) wincl += o; }
            const uint32_t wexcl = wincl - tsum;
            S.tpre[gs & 1][t] = wexcl;
            if (lane == 31) S.wtot[gs & 1][w] = wincl;
            const int any_ev = __syncthreads_or(evm != 0u);
            // every stage before this step's is free now (all threads are past their prefix look-ups): top the ring up
            if (t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_upto(gs + K1_STAGES);
            }
            // cross-warp prefix: the first lanes scan the warp totals
            uint32_t x = lane < K1_WARPS ? S.wtot[gs & 1][lane] : 0u;
#pragma unroll
            for (int d = 1; d < K1_WARPS; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += o; }
            const uint32_t total = __shfl_sync(0xffffffffu, x, K1_WARPS - 1);
            uint32_t wbase = __shfl_sync(0xffffffffu, x, (w - 1u) & 31u);
            if (w == 0) wbase = 0;
            // record starts inside this step: prefix value at the record's first op
            {
                const uint32_t ro = t < nr ? S.roff[t] : 0xffffffffu;
                const uint32_t fb = c * K1_CHUNK;
                const bool mine = ro >= fb && ro < fb + K1_CHUNK && ro < span_hi;
                const uint32_t rel = mine ? ro - fb : 0u, owner = rel / K1_V, wp = owner >> 5;
                uint32_t b2 = __shfl_sync(0xffffffffu, x, (wp - 1u) & 31u);
                if (wp == 0) b2 = 0;
                if (mine) S.pstart[t] = carry + b2 + S.tpre[gs & 1][owner] + buf[rel];
            }
            if (flags & 0x40u) {                                         // rust-htslib panics on an unknown op
#pragma unroll
                for (int k = 0; k < K1_V; k++) if ((v[k] & 15u) > 8u) {
                    const uint32_t i = k1_find_read(S.roff, nr, fp0 + k);
                    if (S.rkeep[i]) report(B.ctrl, ra + i, RANK_CIGAR_OP);
                }
            }
            if (any_ev) staged = k1_stage_events(S, O, tile, ra, nr, staged, evm, v, buf + t * K1_V, fp0, carry + wbase + wexcl, spare);
            carry += total;
        }
        if (staged) k1_flush(S, O, tile, ra, nr, staged, spare);
        __syncthreads();
        for (uint32_t i = t; i <= nr; i += K1_THREADS) if (S.roff[i] >= span_hi) S.pstart[i] = carry;   // trailing empty records + sentinel
        if (t == 0 && S.flushes == 0) B.tile_cnt[tile] = 0;
        __syncthreads();
        if (t < nr) {
            const uint32_t T = S.pstart[t + 1] - S.pstart[t];
            const uint32_t info = S.rkeep[t] ? ((S.rcnt[t] & K1_CNT_MASK) | S.rflags[t]) : 0u;
            B.k1[ra + t] = make_uint2(T, info);
        }
        __syncthreads();                                   // the next tile re-initialises the per-record arrays
    }
    if (trace && t == 0) {
        unsigned long long* d = B.dbg + 4ull * blockIdx.x;
        d[0] = tr_start; d[1] = tr_first; d[2] = gtimer(); d[3] = (unsigned long long)ntile | ((unsigned long long)tr_scanned << 32);
    }
    tr.end();
    // the slab still held in reserve was never used: blank it
    const uint32_t last = __shfl_sync(0xffffffffu, spare, 0);
    __shared__ uint32_t s_last;
    if (t == 0) s_last = last;
    __syncthreads();
    const uint32_t slot = B.prim_slots + s_last + t;
    if (t < B.slab && slot < B.raw_cap) reinterpret_cast<uint4*>(B.raw + slot)[0] = make_uint4(0xffffffffu, 0u, 0u, 0u);
}

// ======================================================================================
// kernel 1a: event screen over the flat CIGAR stream (short-record batches)
//
// I/D >= indel_min are sparse in HiFi-like batches, and total_consume / the prefixes are only ever read for records that
// have such an event (kernel 4b).  So the whole CIGAR stream is first streamed once at full width -- plain coalesced 128-bit
// loads, four per thread in flight, four instructions per op (table look-up, OR, compare, predicated OR), no record
// structure at all -- and all it leaves behind is the list of 512-op steps with an event candidate or an unknown op code in
// them.  Kernel 1b resolves the listed steps to records.  It writes nothing per record: a record has an indel summary only
// if kernel 1b claims it (k4a reads the claim bitmap first).
// (Measured: resolving candidates to records inside this kernel -- a 4-level search per flagged step -- cost 12 of 39 us.)
// ======================================================================================
static constexpr int K1A_THREADS = 256;
static constexpr int K1A_VEC = 4;                      // 128-bit loads per thread per step: a warp step covers 2 KB = 512 ops
static constexpr int K1A_CTAS = 8;                     // CTAs per SM: full occupancy (32 registers) measured faster than fewer warps with deeper prefetch
static constexpr uint32_t K1A_STEP_OPS = 32 * K1A_VEC * 4;

// One warp step: screen the 512 ops held in q; true (warp-uniform) when some lane saw an event candidate or an unknown op.
__device__ __forceinline__ bool k1a_step(uint32_t imin16, uint32_t lut_lo, uint32_t lut_hi, const uint4 (&q)[K1A_VEC])
{
    uint32_t sus = 0, flags = 0;
#pragma unroll
    for (int k = 0; k < K1A_VEC; k++) {
        const uint32_t vv[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t f;
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(f) : "r"(lut_lo), "r"(lut_hi), "r"(vv[j]));
            flags |= f;
            asm("{\n.reg .pred p;\nsetp.ge.u32 p, %1, %2;\n@p or.b32 %0, %0, %3;\n}" : "+r"(sus) : "r"(vv[j]), "r"(imin16), "r"(f));
        }
    }
    return __any_sync(0xffffffffu, ((sus & 2u) | (flags & 0x40u)) != 0u);
}

__global__ void __launch_bounds__(K1A_THREADS, K1A_CTAS) k1a_screen(DevBatch B, DevParams P, unsigned long long n_ops)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nthreads = gridDim.x * K1A_THREADS, gt = blockIdx.x * K1A_THREADS + threadIdx.x;
    griddep_launch();                                  // kernel 1b may be placed; it waits for this grid before it reads anything
    CtaTrace tr(B, 8);
    const uint32_t imin16 = P.indel_min >= (1u << 28) ? 0xffffffffu : (P.indel_min << 4);
    const uint32_t nvec = (uint32_t)((n_ops + 3ull) / 4ull);               // < 2^30 (host); the cigar buffer is padded by 16 bytes
    const uint4* cig = reinterpret_cast<const uint4*>(B.cigar);
    // Warp steps are dealt round robin over all warps of the grid (event-dense genome regions are contiguous in the array).
    // The last vector may hold padding past the last op: a false flag there only costs kernel 1b a look.
    const uint32_t wstep = (nthreads >> 5) * (32 * K1A_VEC);
    uint32_t lut_lo, lut_hi;                                                // the op table, pinned in two registers
    asm volatile("mov.u32 %0, %2;\n\tmov.u32 %1, %3;" : "=r"(lut_lo), "=r"(lut_hi) : "n"(K1_LUT_LO), "n"(K1_LUT_HI));
    // flagged steps are remembered in a per-warp bit mask and appended to the step list 32 iterations at a time (one atomic
    // per warp for a typical batch, at its very end): nothing in the streaming loop waits for memory it does not stream
    const uint32_t gw = gt >> 5, nw = nthreads >> 5;
    auto append = [&](uint32_t mask, uint32_t it0) {
        const uint32_t n = __popc(mask);
        if (!n) return;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&B.ctrl->n_flagged, n);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < n) B.step_list[base + lane] = gw + (it0 + __fns(mask, 0, lane + 1)) * nw;
    };
    uint32_t hitmask = 0, it = 0;
    for (uint32_t v0 = gw * (32 * K1A_VEC); v0 < nvec; v0 += wstep, it++) {
        const uint4* p = cig + v0 + lane;
        uint4 q[K1A_VEC];
        if (v0 + 32 * K1A_VEC <= nvec) {
#pragma unroll
            for (int k = 0; k < K1A_VEC; k++) q[k] = __ldg(p + 32 * k);
        } else {
#pragma unroll
            for (int k = 0; k < K1A_VEC; k++) q[k] = v0 + 32 * k + lane < nvec ? __ldg(p + 32 * k) : make_uint4(0u, 0u, 0u, 0u);
        }
        hitmask |= (k1a_step(imin16, lut_lo, lut_hi, q) ? 1u : 0u) << (it & 31u);
        if ((it & 31u) == 31u) { append(hitmask, it - 31u); hitmask = 0; }
    }
    append(hitmask, it & ~31u);
    tr.end();
}

// largest r in [0, n_reads) with cigar_off[r] <= fp, for fp < cigar_off[n_reads]: a 33-ary search done by the whole warp
// (32 probes per step, 4 steps for a million records)
__device__ __forceinline__ uint32_t k1b_find_read(const unsigned long long* off, uint32_t n_reads, unsigned long long fp)
{
    const uint32_t lane = threadIdx.x & 31;
    uint32_t lo = 0, hi = n_reads;                     // invariant: off[lo] <= fp < off[hi]
    while (hi - lo > 1u) {
        const uint32_t n = hi - lo;
        uint32_t probe; bool valid = true;
        if (n <= 33u) { probe = lo + 1u + lane; valid = probe < hi; }
        else probe = lo + (uint32_t)(((unsigned long long)n * (lane + 1u)) / 33ull);       // strictly increasing, inside (lo, hi)
        const bool le = valid && off[valid ? probe : lo] <= fp;
        const uint32_t k = __popc(__ballot_sync(0xffffffffu, le));                     // sorted offsets: the true lanes are 0..k-1
        const uint32_t below = __shfl_sync(0xffffffffu, probe, (k + 31u) & 31u);
        const uint32_t above = __shfl_sync(0xffffffffu, probe, k & 31u);
        const uint32_t above_ok = __shfl_sync(0xffffffffu, (uint32_t)valid, k & 31u);
        if (k < 32u && above_ok) hi = above;
        if (k) lo = below;
    }
    return lo;
}

// ======================================================================================
// kernel 1b: from flagged steps to records, then one thread per record
//
// k1b_claim: a warp takes a step off kernel 1a's list, finds the records that overlap its 512 ops (one search for the first,
// then consecutive offsets) and claims every kept one exactly once on the device (atomic bit per record: a record can overlap
// several flagged steps).  Claimed records go on one of two lists: short ones (walked by a thread each, below) and long
// ones (kernel 1c: the flat block scan, one record per tile, which is balanced for CIGARs of any length -- in ONT batches
// ~10 % of the records hold an event and those are 10^3..10^5 ops long).
// k1b_walk: one thread per short record, warps full.  left_consume, total_consume, the event sequence and the two merge
// predicates are all thread-local (main.rs:523-600, 612-635, 673-678); no staging, no block scan.  The walk reads aligned
// 128-bit vectors, four in flight; neighbouring threads walk neighbouring records, so the sectors they touch are shared, and
// the stream has just been through L2 for kernel 1a.  The loop body is branch-free (lanes walk different records): an event
// is parked in shared memory by a predicated store, the merge predicates are evaluated on the parked events afterwards, the
// events are written out with one reservation per warp.  A record with more events than parking space joins the long list.
// Raw events go to the atomically allocated region (no tile slices here).
// Both grids are one resident wave (measured: a warp per step, flagged or not, was 7.8k CTAs and 16 us of CTA turnover; and
// walking inside the per-step warps ran at a third of the warp width, since most of a step's records belong to a neighbour).
// ======================================================================================
static constexpr int K1B_THREADS = 256;
static constexpr uint32_t K1B_LONG = 256;              // ops; longer records go to kernel 1c
static constexpr uint32_t K1B_EV = 4;                  // events per record parked in shared memory during the walk
static constexpr int K1B_VEC = 4;                      // 128-bit loads in flight per walking thread

__global__ void __launch_bounds__(K1B_THREADS) k1b_claim(DevBatch B, DevParams P, unsigned long long n_ops, uint32_t use_k1c)
{
    const uint32_t t = threadIdx.x, lane = t & 31;
    griddep_wait();                                    // kernel 1a's step list and zeroed summaries
    griddep_launch();
    CtaTrace tr(B, 9);
    const uint32_t n_list = B.ctrl->n_flagged, nw = (gridDim.x * K1B_THREADS) >> 5;
    for (uint32_t li = (blockIdx.x * K1B_THREADS + t) >> 5; li < n_list; li += nw) {
        const bool trace = B.dbg && B.dbg_sel == 1u;
        const unsigned long long tr0 = trace ? gtimer() : 0ull;
        unsigned long long tr1 = 0;
        const uint32_t st = B.step_list[li];
        const unsigned long long lo_op = (unsigned long long)st * K1A_STEP_OPS, hi_op = min(lo_op + K1A_STEP_OPS, n_ops);
        if (lo_op >= hi_op) continue;
        const uint32_t r_first = k1b_find_read(B.cigar_off, B.n_reads, lo_op);
        if (trace) tr1 = gtimer();
        for (uint32_t rb = r_first;; rb += 32) {                              // records overlapping the step, 32 at a time
            const uint32_t r = rb + lane;
            unsigned long long o0 = 0, o1 = 0;
            uint32_t flag = 0, mapq = 0;
            const bool in = r < B.n_reads;
            if (in) { o0 = B.cigar_off[r]; o1 = B.cigar_off[r + 1]; flag = B.flag[r]; mapq = B.mapq[r]; }
            const bool overlaps = in && o0 < hi_op;                           // (o1 > lo_op holds from r_first on, empty records aside)
            bool mine = overlaps && o1 > o0 && o1 > lo_op && keep_record(P, flag, mapq);
            {   // the 32 records of the warp share at most two words of the claim bitmap: two atomics per warp, not one per lane
                const uint32_t cb = __ballot_sync(0xffffffffu, mine), sh = rb & 31u;
                const uint32_t lo_bits = cb << sh, hi_bits = sh ? cb >> (32u - sh) : 0u;
                uint32_t old = 0;
                if (lane == 0 && lo_bits) old = atomicOr(&B.dirty_bits[rb >> 5], lo_bits);
                if (lane == 1 && hi_bits) old = atomicOr(&B.dirty_bits[(rb >> 5) + 1u], hi_bits);
                const uint32_t old_lo = __shfl_sync(0xffffffffu, old, 0), old_hi = __shfl_sync(0xffffffffu, old, 1);
                const uint32_t p = sh + lane;
                if ((p < 32u ? old_lo >> p : old_hi >> (p - 32u)) & 1u) mine = false;
            }
            // long records: kernel 1c's list, or (batches of short records, where kernel 1c is not launched) the short list with bit 31
            // set: k1b_walk then scans them with a whole warp
            const bool is_big = mine && o1 - o0 > K1B_LONG, is_long = is_big && use_k1c;
            const uint32_t sm = __ballot_sync(0xffffffffu, mine && !is_long), lm = __ballot_sync(0xffffffffu, is_long);
            uint32_t sbase = 0, lbase = 0;
            if (lane == 0 && sm) sbase = atomicAdd(&B.ctrl->n_short, (uint32_t)__popc(sm));
            if (lane == 1 && lm) lbase = atomicAdd(&B.ctrl->n_long, (uint32_t)__popc(lm));
            sbase = __shfl_sync(0xffffffffu, sbase, 0); lbase = __shfl_sync(0xffffffffu, lbase, 1);
            const uint32_t below = (1u << lane) - 1u;
            if (mine && !is_long) B.short_list[sbase + __popc(sm & below)] = r | (is_big ? 0x80000000u : 0u);
            if (is_long) B.long_list[lbase + __popc(lm & below)] = r;
            // the next 32 records matter only if the last one of these still ends inside the step
            if (!__shfl_sync(0xffffffffu, (uint32_t)(overlaps && o1 < hi_op), 31)) break;
        }
        if (trace && lane == 0) {                                             // EXLR_OPT_TRACE: {start, search done, end, 0}
            unsigned long long* d = B.dbg + 4ull * (li & 8191u);
            d[0] = tr0; d[1] = tr1; d[2] = gtimer(); d[3] = 0;
        }
    }
    tr.end();
}

__global__ void __launch_bounds__(K1B_THREADS) k1b_walk(DevBatch B, DevParams P, uint32_t use_k1c)
{
    __shared__ uint2 s_ev[K1B_THREADS][K1B_EV + 1];    // {left_consume, op word}; the extra slot tells "more than K1B_EV"
    const uint32_t t = threadIdx.x, lane = t & 31;
    griddep_wait();                                    // k1b_claim's lists
    griddep_launch();
    CtaTrace tr(B, 10);
    const uint32_t n_list = B.ctrl->n_short, stride = gridDim.x * K1B_THREADS;
    const uint32_t imin16 = P.indel_min >= (1u << 28) ? 0xffffffffu : (P.indel_min << 4);
    const uint4* cig4 = reinterpret_cast<const uint4*>(B.cigar);
    const uint32_t ev0 = smem_u32(&s_ev[t][0]), ev_end = ev0 + (K1B_EV + 1) * 8u;
    for (uint32_t i0 = blockIdx.x * K1B_THREADS + (t & ~31u); i0 < n_list; i0 += stride) {     // warp-uniform trip count
        const uint32_t i = i0 + lane;
        const bool have = i < n_list;
        uint32_t r = 0, pos2 = 0, nv = 0, head = 0, nops = 0;
        const uint4* c4 = cig4;
        bool is_long = false;                                                  // not for one thread: too long, or too many events
        if (have) { r = B.short_list[i]; is_long = r >> 31; r &= 0x7fffffffu; }
        if (have && !is_long) {
            const unsigned long long o0 = B.cigar_off[r], o1 = B.cigar_off[r + 1], a0 = o0 & ~3ull;
            pos2 = (uint32_t)B.pos[r];
            nv = (uint32_t)((((o1 + 3ull) & ~3ull) - a0) >> 2);                // aligned vectors spanned (<= 65)
            head = (uint32_t)(o0 - a0); nops = (uint32_t)(o1 - o0);
            c4 = cig4 + (a0 >> 2);
        }
        uint32_t L = 0, flags = 0, evp = ev0;
        for (uint32_t vb = 0; vb < nv; vb += K1B_VEC) {
            uint4 q[K1B_VEC];
#pragma unroll
            for (int k = 0; k < K1B_VEC; k++) q[k] = vb + k < nv ? __ldg(c4 + vb + k) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int k = 0; k < K1B_VEC; k++) {
                const uint32_t vv[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t e = (vb + k) * 4u + j;                     // element index counted from the aligned base
                    const uint32_t v = (e - head) < nops ? vv[j] : 0u;         // outside the record: reads as 0M
                    uint32_t f;
                    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(f) : "r"(K1_LUT_LO), "r"(K1_LUT_HI), "r"(v));
                    flags |= f;
                    // an event (I/D >= indel_min, main.rs:553,569) parks {left_consume, op} and bumps the parking pointer, all predicated
                    asm volatile("{\n.reg .pred p, q;\n"
                                 "setp.ne.u32 q, %3, 0;\n"
                                 "setp.ge.and.u32 p, %2, %4, q;\n"
                                 "setp.lt.and.u32 p, %0, %5, p;\n"
                                 "@p st.shared.v2.u32 [%0], {%1, %2};\n"
                                 "@p add.u32 %0, %0, 8;\n}"
                                 : "+r"(evp) : "r"(L), "r"(v), "r"(f & 2u), "r"(imin16), "r"(ev_end) : "memory");
                    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(L) : "r"(f & 1u), "r"(v >> 4));   // M D N = consume the reference (main.rs:528-545)
                }
            }
        }
        const uint32_t cnt = (evp - ev0) >> 3;
        if (have && !is_long) {
            if (flags & 0x40u) report(B.ctrl, r, RANK_CIGAR_OP);                                    // rust-htslib panics on an unknown op
            if (cnt > K1B_EV) is_long = true;                                                       // more events than parking space
            else {
                uint32_t info = 0;
                for (uint32_t j = 1; j < cnt; j++) {
                    const uint2 x = s_ev[t][j - 1], y = s_ev[t][j];
                    if ((x.y & 15u) == 2u && (y.y & 15u) == 2u) {
                        if (j == 1u && abs_diff(pos2 + y.x, pos2 + x.x + (x.y >> 4)) < P.merge_min) info |= K1_PAIR_MERGE;   // main.rs:615
                        if (abs_diff(pos2 + x.x, pos2 + y.x + (y.y >> 4)) < P.merge_min) info |= K1_FAR_HIT;                  // main.rs:673-678
                    }
                }
                B.k1[r] = make_uint2(L, (cnt & K1_CNT_MASK) | info);
            }
        }
        // parked events: warp scan of the counts, one reservation, every lane writes its own
        const uint32_t parked = have && !is_long ? cnt : 0u;
        uint32_t incl = parked;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t base = 0;
        if (lane == 0 && total) base = atomicAdd(&B.ctrl->n_raw, total);
        base = B.prim_slots + __shfl_sync(0xffffffffu, base, 0) + incl - parked;
        for (uint32_t j = 0; j < parked; j++) {
            if (base + j < B.raw_cap) {
                const uint2 e = s_ev[t][j];
                uint4* d = reinterpret_cast<uint4*>(B.raw + base + j);
                d[0] = make_uint4(r, j, e.x, (e.y >> 4) | ((e.y & 15u) == 2u ? 0x80000000u : 0u));
                d[1] = make_uint4(j ? s_ev[t][j - 1].x : 0u, 0u, 0u, 0u);
            } else B.ctrl->overflow = 1;
        }
        uint32_t lm = __ballot_sync(0xffffffffu, is_long);
        if (lm && use_k1c) {                                                   // on to kernel 1c's list
            uint32_t lbase = 0;
            if (lane == 0) lbase = atomicAdd(&B.ctrl->n_long, (uint32_t)__popc(lm));
            lbase = __shfl_sync(0xffffffffu, lbase, 0);
            if (is_long) B.long_list[lbase + __popc(lm & ((1u << lane) - 1u))] = r;
        } else {
            for (; lm; lm &= lm - 1u) k1_warp_record(B, P, __shfl_sync(0xffffffffu, r, __ffs((int)lm) - 1));   // the whole warp scans it
        }
    }
    tr.end();
}

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
static constexpr int K3B_THREADS = 192;                 // the usual tile (64 records, ~140 pieces + 64 own segments) parses in one round
static constexpr int K3B_TILE = 64;                     // SA records per tile (pieces are then spread over all threads)
static constexpr int K3B_CTAS = 7;                      // CTAs per SM: 30 KB of shared memory and 48 registers x 192 threads each
static constexpr uint32_t K3B_STAGE_BYTES = 12 * 1024;
static constexpr uint32_t K3B_SEMI = 8;                // ';' positions kept per record by phase 1; more -> phase 1b rescans
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

// Fast path of parse_supplementary_alignment (utils.rs:119-139) for the regular form every aligner writes,
//     chrom,<1-9 digits>,<+|->,(<1-9 digits><op>)+,<1-3 digits <= 255>,<1-9 digits>
// in ONE pass over the bytes.  Anything else -- signs, longer numbers, missing or extra fields, odd bytes, trailing digits in
// the CIGAR -- returns false and the caller runs the exact dev_parse_piece above, which also yields the reference's panics.
// For the accepted form both produce the same Seg: same field boundaries, u32-wrapping sums, u64 key.
// Written for warp convergence: one simple loop per field, no early exit (a failed check only clears `ok`), and the lanes
// of `m` (the lanes of the warp that hold a piece) re-join after every loop -- with early returns and nested digit loops
// the lanes drifted apart and the parser ran at a quarter of the warp width.
__device__ __forceinline__ bool dev_parse_piece_fast(const uint8_t* p /* staged bytes */, uint32_t bias /* sa offset of p[0] */,
                                                     uint32_t b, uint32_t e, const DevParams& P, Seg* out, uint32_t m)
{
    b -= bias; e -= bias;
    bool ok = true;
    uint32_t i = b;
    while (i < e && p[i] != ',') i++;                                          // chrom
    __syncwarp(m);
    const uint32_t ce = i;
    ok &= i < e;
    i++;
    uint32_t pos = 0, nd = 0;                                                  // pos
    while (i < e) { const uint32_t d = (uint32_t)p[i] - '0'; if (d > 9u) break; pos = pos * 10u + d; nd++; i++; }
    __syncwarp(m);
    ok &= nd - 1u < 9u && i < e && p[min(i, e - 1u)] == ',';
    i++;
    const uint32_t sc = i < e ? (uint32_t)p[i] : 0u, sc2 = i + 1u < e ? (uint32_t)p[i + 1] : 0u;   // strand
    ok &= (sc == '+' || sc == '-') && sc2 == ',';
    i += 2;
    // CIGAR text (utils.rs:88-117, 12-42): the same step for every byte up to the comma
    uint32_t sS = 0, sH = 0, sD = 0, sM = 0, sE = 0, sX = 0, nops = 0, n = 0, bad = 0;
    unsigned long long key = 0; bool seenM = false;
    nd = 0;
    while (i < e) {
        const uint32_t c = p[i];
        if (c == ',') break;
        const uint32_t d = c - '0';
        if (d <= 9u) { n = n * 10u + d; nd++; }
        else {
            const uint32_t x = c - '=';                                         // = D H I M N P S X  ->  0 7 11 12 16 17 19 22 27
            bad |= (x >= 28u || !((0x84B1881u >> (x & 31u)) & 1u) || nd - 1u >= 9u) ? 1u : 0u;
            sM += c == 'M' ? n : 0u; sS += c == 'S' ? n : 0u; sD += c == 'D' ? n : 0u;
            sH += c == 'H' ? n : 0u; sE += c == '=' ? n : 0u; sX += c == 'X' ? n : 0u;
            if (!seenM && x < 28u && ((0x8401001u >> x) & 1u)) key += n;        // = I S X before the first M (utils.rs:33)
            seenM |= c == 'M';
            n = 0; nd = 0; nops++;
        }
        i++;
    }
    __syncwarp(m);
    ok &= !bad && nd == 0u && nops != 0u && i < e;                             // ends at the comma, right after an op
    i++;
    uint32_t mq = 0;                                                           // mapq: u8
    nd = 0;
    while (i < e) { const uint32_t d = (uint32_t)p[i] - '0'; if (d > 9u) break; mq = mq * 10u + d; nd++; i++; }
    __syncwarp(m);
    ok &= nd - 1u < 3u && mq <= 255u && i < e && p[min(i, e - 1u)] == ',';
    i++;
    nd = 0;                                                                    // NM: parsed, value unused (utils.rs:135)
    while (i < e) { if ((uint32_t)p[i] - '0' > 9u) break; nd++; i++; }
    __syncwarp(m);
    ok &= nd - 1u < 9u && i == e;
    if (!ok) return false;
    uint32_t cb = b;
    if (ce - cb >= 3u && p[cb] == 'c' && p[cb + 1] == 'h' && p[cb + 2] == 'r') cb += 3;
    out->chrom_ref = 0x80000000u | (cb + bias);
    out->chrom_len = ce - cb;
    out->start = (int64_t)pos - 1;
    out->end = out->start + (int64_t)sD + (int64_t)sM + (int64_t)sE + (int64_t)sX;
    out->key = (int64_t)key;
    out->clip_big = (sS > P.ins_clip_min || sH > P.ins_clip_min) ? 1u : 0u;
    out->strand_neg = sc == '-';
    return true;
}

// Calls f(w0, z) for every aligned word of bytes [b0, e0): z has 0x80 in each byte that equals the byte replicated in c4,
// bytes outside the range masked off (only the first and the last word pay for the masking).
template <class F>
__device__ __forceinline__ void swar_scan(const SmemBytes& s, uint32_t b0, uint32_t e0, uint32_t c4, F f)
{
    if (b0 >= e0) return;
    uint32_t w0 = b0 & ~3u;
    const uint32_t last = (e0 - 1u) & ~3u;
    uint32_t z = swar_eq(s.word(w0), c4) & (0xffffffffu << (8u * (b0 - w0)));
    if (w0 == last) { f(w0, z & (0xffffffffu >> (8u * (w0 + 4u - e0)))); return; }
    f(w0, z);
    for (w0 += 4u; w0 < last; w0 += 4u) f(w0, swar_eq(s.word(w0), c4));
    f(last, swar_eq(s.word(last), c4) & (0xffffffffu >> (8u * (last + 4u - e0))));
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
__device__ __forceinline__ void k3b_record_seg(const DevBatch& B, const DevParams& P, uint32_t j, uint32_t r, Seg* out)
{
    const SaSum sum = B.sa_sum[j];
    const int32_t tid = B.tid[r];
    out->chrom_ref = (uint32_t)tid; out->chrom_len = B.ref_off[tid + 1] - B.ref_off[tid];
    out->start = (int64_t)B.pos[r]; out->end = out->start + sum.refspan; out->key = sum.ffm;
    out->clip_big = (sum.S > P.ins_clip_min || sum.H > P.ins_clip_min) ? 1u : 0u;
    out->strand_neg = (B.flag[r] & 0x10u) ? 1u : 0u;
}

// Fallback: one SA record parsed start to finish by one thread (used when a tile's SA bytes or segment count do
// not fit the staged layout, e.g. -k far above the default).  Must be called by every lane of the warp.
template <class Bytes>
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
            k3b_record_seg(B, P, j, r, &segs[0]);
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

// Fast path layout: the tile's SA bytes, its piece list and its segments all live in shared memory, and the work
// is re-flattened between phases so that the lanes of a warp always run the same loop:
//   phase 1  thread per record : count ';' (the -k cap) and non-empty pieces          -> slots per record, block scan
//   phase 1b thread per record : write the [begin,end) of every piece into the piece list
//   phase 2  thread per PIECE  : parse_supplementary_alignment on homogeneous pieces   -> Seg
//   phase 3  thread per record : record segment, sort, rules, emit
struct __align__(16) K3bSmem {
    uint8_t bytes[K3B_STAGE_BYTES];
    Seg segs[K3B_MAXP];
    uint32_t pb[K3B_MAXP], pe[K3B_MAXP];
    uint32_t rerr[K3B_TILE];
    uint32_t semi[K3B_TILE][K3B_SEMI];                 // positions of the first ';' of every record (phase 1 -> 1b)
    uint32_t wsum[K3B_THREADS / 32];
    uint8_t pread[K3B_MAXP];
};

__global__ void __launch_bounds__(K3B_THREADS, K3B_CTAS) k3b_sa_events(DevBatch B, DevParams P)
{
    extern __shared__ __align__(16) unsigned char k3b_smem_raw[];
    K3bSmem& S = *reinterpret_cast<K3bSmem*>(k3b_smem_raw);
    griddep_wait();                                    // kernel 3a's summaries
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
            k3b_record(B, P, s, j, active, local_segs);
            continue;
        }
        for (uint32_t o = a0 + t * 16u; o < span_e; o += K3B_THREADS * 16u)
            *reinterpret_cast<uint4*>(S.bytes + (o - a0)) = __ldg(reinterpret_cast<const uint4*>(B.sa_bytes + o));
        // phase 1: pieces per record
        uint32_t r = 0, b0 = 0, e0 = 0, slots = 0;
        bool dropped = false, is_str = false;
        if (active) { r = B.sa_list[j]; b0 = B.sa_off[r]; e0 = B.sa_off[r + 1]; is_str = B.sa_kind[r] == EXLR_SA_STRING; }
        __syncthreads();
        SmemBytes s{S.bytes, a0};
        uint32_t nsemi = 0;
        if (active) {
            uint32_t nonempty = 0;
            if (is_str) {
                // one pass, four bytes per step: every ';' closes a piece (empty when it directly follows the previous ';' or the
                // start of the string); the last piece runs to the end of the string.  pieces = #';' + 1 (main.rs:309).
                uint32_t pbeg = b0;
                swar_scan(s, b0, e0, 0x3b3b3b3bu, [&](uint32_t w0, uint32_t z) {
                    while (z) {
                        const uint32_t i = w0 + ((uint32_t)(__ffs((int)z) - 1) >> 3);
                        z &= z - 1u;
                        if (nsemi < K3B_SEMI) S.semi[t][nsemi] = i;
                        nsemi++;
                        nonempty += i > pbeg;
                        pbeg = i + 1u;
                    }
                });
                nonempty += e0 > pbeg;
                if ((unsigned long long)nsemi + 1ull > P.max_supp_alignm) dropped = true;   // main.rs:311-313
            }
            slots = dropped ? 0u : 1u + nonempty;
            S.rerr[t] = 0xffffffffu;
        }
        uint32_t incl = slots;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
        if (lane == 31) S.wsum[w] = incl;
        __syncthreads();
        uint32_t sb = incl - slots, total = 0;
#pragma unroll
        for (int k = 0; k < K3B_THREADS / 32; k++) { const uint32_t x = S.wsum[k]; if ((uint32_t)k < w) sb += x; total += x; }
        if (total > K3B_MAXP) {                                               // block-uniform: too many segments for the staged layout
            k3b_record(B, P, s, j, active, local_segs);
            continue;
        }
        // phase 1b: piece list
        if (active && !dropped) {
            S.pb[sb] = 0xffffffffu; S.pread[sb] = (uint8_t)t;                 // slot 0 of the record: its own alignment
            if (is_str) {
                uint32_t at = sb + 1, pbeg = b0;
                auto piece_end = [&](uint32_t i) {
                    if (i > pbeg) { S.pb[at] = pbeg; S.pe[at] = i; S.pread[at] = (uint8_t)t; at++; }   // filter(|x| x.len() > 0), main.rs:315
                    pbeg = i + 1u;
                };
                if (nsemi <= K3B_SEMI) {
                    for (uint32_t k = 0; k < nsemi; k++) piece_end(S.semi[t][k]);
                } else {                                                       // more ';' than phase 1 kept (large -k): rescan
                    swar_scan(s, b0, e0, 0x3b3b3b3bu, [&](uint32_t w0, uint32_t z) {
                        while (z) { piece_end(w0 + ((uint32_t)(__ffs((int)z) - 1) >> 3)); z &= z - 1u; }
                    });
                }
                if (e0 > pbeg) piece_end(e0);
            }
        }
        __syncthreads();
        // phase 2: one thread per piece
        for (uint32_t x0 = 0; x0 < total; x0 += K3B_THREADS) {               // block-uniform trip count
            const uint32_t x = x0 + t;
            const uint32_t pbeg = x < total ? S.pb[x] : 0xffffffffu;
            const bool has = pbeg != 0xffffffffu;
            const uint32_t m = __ballot_sync(0xffffffffu, has);
            if (!has) continue;
            const uint32_t pend = S.pe[x];
            if (dev_parse_piece_fast(S.bytes, a0, pbeg, pend, P, &S.segs[x], m)) continue;
            const uint32_t err = dev_parse_piece(s, pbeg, pend, P, &S.segs[x]);   // irregular piece: the exact parser decides
            if (err) atomicMin(&S.rerr[S.pread[x]], (x << 8) | err);          // the first failing piece in SA order wins
        }
        __syncthreads();
        tr.mid();
        // phase 3: one thread per record
        uint32_t nseg = 0;
        if (active && !dropped) {
            uint32_t err = S.rerr[t];
            nseg = slots;
            if (err != 0xffffffffu) { report(B.ctrl, r, err & 0xffu); nseg = 0; }
            else k3b_record_seg(B, P, j, r, &S.segs[sb]);
        }
        k3b_finish(B, P, s, j, active, r, dropped, &S.segs[sb < K3B_MAXP ? sb : 0], nseg);
    }
    tr.end();
}

// ======================================================================================
// kernel 4a: per-record line count (with the pair merge) -> chained scan -> line offsets
// ======================================================================================
__device__ __forceinline__ uint32_t indel_lines(uint32_t info)
{
    const uint32_t cnt = info & K1_CNT_MASK;
    return (cnt == 2u && (info & K1_PAIR_MERGE)) ? 1u : cnt;                  // main.rs:612-635
}

__device__ __forceinline__ uint32_t record_lines(const DevBatch& B, uint32_t r, uint32_t csa, uint32_t info)
{
    if (csa & CSA_DROP) return 0u;                                            // -k cap: no lines at all (main.rs:311-313)
    if ((info & K1_CNT_MASK) > 2u && (info & K1_FAR_HIT)) report(B.ctrl, r, RANK_MERGE_DOMAIN);
    return (csa & CSA_CNT_MASK) + indel_lines(info);
}

__global__ void __launch_bounds__(SCAN_THREADS) k4a_line_scan(DevBatch B, DevParams P)
{
    __shared__ uint32_t s_tile, s_warp[16];
    griddep_launch();                                  // kernel 4b may be placed; it waits for this grid before it reads anything
    CtaTrace tr(B, 5);
    if (threadIdx.x == 0) s_tile = atomicAdd(&B.ctrl->ticket_b, 1u);
    __syncthreads();
    const uint32_t tile = s_tile, n = B.n_reads;
    const uint32_t r0 = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t c[SCAN_ITEMS], mine = 0;
    const bool full = r0 + SCAN_ITEMS <= n;
    if (full && B.k1_gated) {
        // screened CIGAR path: a record has an indel summary only if kernel 1b claimed it (its bit in the claim bitmap);
        // everything else is "no event" without ever having been written -- 16 records share one 16-bit slice of the bitmap
        union { uint4 v[4]; uint32_t u[16]; } cs;
#pragma unroll
        for (int k = 0; k < 4; k++) cs.v[k] = reinterpret_cast<const uint4*>(B.csa + r0)[k];
        const uint32_t bits = (B.dirty_bits[r0 >> 5] >> (r0 & 31u)) & 0xffffu;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            const uint32_t info = (bits >> i) & 1u ? B.k1[r0 + i].y : 0u;
            c[i] = record_lines(B, r0 + i, cs.u[i], info); mine += c[i];
        }
    } else if (full) {
        union { uint4 v[4]; uint32_t u[16]; } cs;
        union { uint4 v[8]; uint2 p[16]; } k1;
#pragma unroll
        for (int k = 0; k < 4; k++) cs.v[k] = reinterpret_cast<const uint4*>(B.csa + r0)[k];
#pragma unroll
        for (int k = 0; k < 8; k++) k1.v[k] = reinterpret_cast<const uint4*>(B.k1 + r0)[k];
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { c[i] = record_lines(B, r0 + i, cs.u[i], k1.p[i].y); mine += c[i]; }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            const uint32_t r = r0 + i;
            uint32_t info = 0;
            if (r < n && (!B.k1_gated || ((B.dirty_bits[r >> 5] >> (r & 31u)) & 1u))) info = B.k1[r].y;
            c[i] = r < n ? record_lines(B, r, B.csa[r], info) : 0u;
            mine += c[i];
        }
    }
    uint32_t grand;
    tr.mid();
    uint32_t at = tile_excl_scan(B.scan_b, tile, mine, s_warp, &grand);
    if (full) {
        union { uint4 v[4]; uint32_t u[16]; } o;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { o.u[i] = at; at += c[i]; }
#pragma unroll
        for (int k = 0; k < 4; k++) reinterpret_cast<uint4*>(B.line_off + r0)[k] = o.v[k];
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) { if (r0 + i < n) B.line_off[r0 + i] = at; at += c[i]; }
    }
    if (threadIdx.x == 0 && tile == (n + SCAN_TILE - 1) / SCAN_TILE - 1) {
        B.line_off[n] = grand;
        B.ctrl->n_events = grand;
        if (grand > B.max_events) B.ctrl->overflow = 1;
    }
    tr.end();
}


// ======================================================================================
// kernel 4b: ordered compaction into the output event buffer
// ======================================================================================
// one raw indel event -> its output line (AlignmentEvent::new, aligments_event.rs:28-57; pair merge main.rs:612-635)
__device__ __forceinline__ void k4b_indel(const DevBatch& B, const RawEv* src)
{
    const uint4 a = reinterpret_cast<const uint4*>(src)[0];
    const uint32_t r = a.x;
    if (r == 0xffffffffu) return;
    const uint32_t csa = B.csa[r];
    if (csa & CSA_DROP) return;
    const uint2 k1 = B.k1[r];
    const uint32_t cnt = k1.y & K1_CNT_MASK;
    const bool merged = cnt == 2u && (k1.y & K1_PAIR_MERGE);
    uint32_t seq = a.y;
    const uint32_t L = a.z, len = a.w & 0x7fffffffu, del = a.w >> 31;
    const uint32_t pos2 = (uint32_t)B.pos[r];
    uint32_t ls = pos2, le, rs, re;
    if (merged) {
        if (seq == 0) return;
        const uint32_t prevL = reinterpret_cast<const uint4*>(src)[1].x;
        le = pos2 + prevL; rs = pos2 + L + len; re = pos2 + k1.x; seq = 0;
    } else if (del) { le = pos2 + L; rs = pos2 + L + len; re = pos2 + k1.x; }      // rend = pos + total_consume
    else { le = pos2 + L; rs = pos2 + L; re = pos2 + L + len; }                    // Ins: length on the right (main.rs:570-577)
    const uint32_t dst = B.line_off[r] + (csa & CSA_CNT_MASK) + seq;
    if (dst >= B.max_events) return;
    const uint32_t neg = (B.flag[r] >> 4) & 1u, tid = (uint32_t)B.tid[r];
    store_event(B.events + dst, (int64_t)ls, (int64_t)le, (int64_t)rs, (int64_t)re, r, tid, tid,
                EXLR_EV_META(1u, EXLR_KIND_INDEL, neg, neg));
}

// One thread per raw slot / overflow entry / SA record.  The overflow and SA-record counts live on the device, so the grid
// is one resident wave striding over the largest of the three ranges (measured: a grid sized for the worst case launched
// 7-9k CTAs, most of them empty, and the launch alone took ~15 us).
__global__ void __launch_bounds__(256) k4b_place(DevBatch B, DevParams P)
{
    griddep_wait();                                    // kernel 4a's line offsets
    griddep_launch();
    CtaTrace tr(B, 6);
    const uint32_t room = B.raw_cap - B.prim_slots, n_ovf = min(B.ctrl->n_raw, room), n_sa = B.ctrl->n_sa;
    tr.mid();
    const uint32_t limit = max(B.prim_slots, max(n_ovf, n_sa)), stride = gridDim.x * blockDim.x;
    for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < limit; x += stride) {
        // indel events, per-tile slices first (slice of tile i = raw[i << capt_log2 ..], tile_cnt[i] entries used) ...
        if (x < B.prim_slots && (x & ((1u << B.capt_log2) - 1u)) < B.tile_cnt[x >> B.capt_log2]) k4b_indel(B, B.raw + x);
        // ... then the shared overflow region
        if (x < n_ovf) k4b_indel(B, B.raw + B.prim_slots + x);
        // SA-derived events: per record contiguous in the temp buffer, they lead the record's lines
        if (x < n_sa) {
            const uint32_t j = x, r = B.sa_list[j];
            const uint32_t csa = B.csa[r];
            if (csa & CSA_DROP) continue;
            const uint32_t cnt = csa & CSA_CNT_MASK, base = B.sa_base[j], dst = B.line_off[r];
            if ((unsigned long long)dst + cnt > B.max_events || (unsigned long long)base + cnt > B.max_events) continue;
            const uint4* src = reinterpret_cast<const uint4*>(B.sa_ev + base);
            uint4* d = reinterpret_cast<uint4*>(B.events + dst);
            for (uint32_t k = 0; k < cnt * 3; k++) d[k] = src[k];
        }
    }
    tr.end();
}

// ======================================================================================
// kernels 5a / 5b: the output lines themselves (optional; EXLR_OPT_DEVICE_FORMAT)
//
// get_alignment_event_record / get_alignment_split_record without -v (utils.rs:225-236, 269-280):
//   lchrom \t lstart \t lend \t lstrand \t rchrom \t rstart \t rend \t rstrand \t events_num \n
// 5a: one thread per event computes the byte length of its line; a chained scan turns the lengths into byte offsets.
// 5b: one thread per event writes its line at its offset -- the lines of consecutive events are consecutive in memory, so
// the byte stores of a warp fall into the same few sectors and merge in L2.
// With this the device-to-host copy carries the final bytes and the host formatter (the slowest host stage after BGZF
// inflate) is not needed; -v lines carry the read name, which never travels to the device: those stay with the host.
// ======================================================================================
__device__ __forceinline__ uint32_t dec_len(int64_t v)
{
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    uint32_t n = v < 0 ? 2u : 1u;
    while (u >= 10ull) { u /= 10ull; n++; }
    return n;
}

__device__ __forceinline__ uint8_t* put_dec(uint8_t* p, int64_t v)
{
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    uint8_t tmp[20]; int n = 0;
    do { tmp[n++] = (uint8_t)('0' + (uint32_t)(u % 10ull)); u /= 10ull; } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

// name of a chrom reference (exlr.h): header name by tid, or the bytes of the SA string up to the next ','
__device__ __forceinline__ const uint8_t* chrom_name(const DevBatch& B, uint32_t ref, uint32_t* len)
{
    if (ref >> 31) {
        const uint8_t* p = B.sa_bytes + (ref & 0x7fffffffu);
        uint32_t n = 0;
        while (p[n] != ',') n++;
        *len = n;
        return p;
    }
    const uint32_t a = B.ref_off[ref];
    *len = B.ref_off[ref + 1] - a;
    return B.ref_bytes + a;
}

__global__ void __launch_bounds__(SCAN_THREADS) k5a_line_bytes(DevBatch B)
{
    __shared__ uint32_t s_tile, s_warp[16];
    griddep_wait();                                    // kernel 4b's events
    griddep_launch();
    CtaTrace tr(B, 12);
    const uint32_t n = min(B.ctrl->n_events, B.max_events), n_tiles = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&B.ctrl->ticket_c, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;
        const uint32_t i = tile * SCAN_THREADS + threadIdx.x;
        uint32_t len = 0;
        if (i < n) {
            const exlr_event e = B.events[i];
            uint32_t a, b;
            chrom_name(B, e.lchrom, &a); chrom_name(B, e.rchrom, &b);
            len = a + b + dec_len(e.lstart) + dec_len(e.lend) + dec_len(e.rstart) + dec_len(e.rend)
                  + (EXLR_EV_LSTRAND(e.meta) < 0 ? 2u : 1u) + (EXLR_EV_RSTRAND(e.meta) < 0 ? 2u : 1u)
                  + dec_len((int64_t)EXLR_EV_NUM(e.meta)) + 9u;                 // 8 tabs and the newline
        }
        uint32_t grand;
        const uint32_t at = tile_excl_scan(B.scan_c, tile, len, s_warp, &grand);
        if (i < n) B.text_off[i] = at;
        if (threadIdx.x == 0 && tile == n_tiles - 1) { B.text_off[n] = grand; B.ctrl->text_bytes = grand; }
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) { B.text_off[0] = 0; B.ctrl->text_bytes = 0; }
    tr.end();
}

// one line into p (any address space); returns the end
__device__ __forceinline__ uint8_t* k5_put_line(const DevBatch& B, const exlr_event& e, uint8_t* p)
{
    uint32_t len;
    const uint8_t* nm = chrom_name(B, e.lchrom, &len);
    for (uint32_t k = 0; k < len; k++) *p++ = nm[k];
    *p++ = '\t'; p = put_dec(p, e.lstart); *p++ = '\t'; p = put_dec(p, e.lend); *p++ = '\t';
    p = put_dec(p, EXLR_EV_LSTRAND(e.meta)); *p++ = '\t';
    nm = chrom_name(B, e.rchrom, &len);
    for (uint32_t k = 0; k < len; k++) *p++ = nm[k];
    *p++ = '\t'; p = put_dec(p, e.rstart); *p++ = '\t'; p = put_dec(p, e.rend); *p++ = '\t';
    p = put_dec(p, EXLR_EV_RSTRAND(e.meta)); *p++ = '\t';
    p = put_dec(p, (int64_t)EXLR_EV_NUM(e.meta)); *p++ = '\n';
    return p;
}

// A warp takes 32 consecutive events: their lines are one contiguous byte range of the output (~1.6 KB).  Every lane formats
// its line into the warp's shared-memory stage at the line's offset inside that range; the stage is laid out with the same
// alignment (mod 16) as the destination, so the warp then writes the range out with coalesced 128-bit stores (bytes only at
// the two ragged ends).  Measured: one thread writing its line straight to global memory, byte by byte, took 92 us for 268k lines.
static constexpr uint32_t K5B_STAGE = 4096;             // bytes of stage per warp (a range that does not fit goes out byte-wise)

__global__ void __launch_bounds__(256) k5b_format(DevBatch B)
{
    __shared__ __align__(16) uint8_t s_stage[8][K5B_STAGE];
    griddep_wait();                                    // kernel 5a's offsets
    CtaTrace tr(B, 13);
    const uint32_t n = min(B.ctrl->n_events, B.max_events), lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    if (B.ctrl->text_bytes > B.text_cap) return;        // the host falls back to its own formatter
    for (uint32_t i0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; i0 < n; i0 += nw * 32u) {
        const uint32_t i = i0 + lane, i1 = min(i0 + 32u, n);
        const uint32_t lo = B.text_off[i0], hi = B.text_off[i1];          // the warp's byte range
        const uint32_t skew = lo & 15u;
        const bool have = i < n;
        exlr_event e;
        uint32_t off = 0;
        if (have) { e = B.events[i]; off = B.text_off[i]; }
        if (hi - lo + skew <= K5B_STAGE) {
            uint8_t* st = s_stage[w];
            if (have) k5_put_line(B, e, st + skew + (off - lo));
            __syncwarp();
            uint8_t* dst = B.text + (lo - skew);                           // 16-byte aligned; stage byte k <-> dst byte k
            const uint32_t end = skew + (hi - lo);
            const uint32_t v0 = skew ? 16u : 0u, v1 = end & ~15u;           // [v0, v1) is whole vectors
            if (v1 > v0) {
                for (uint32_t k = v0 + lane * 16u; k < v1; k += 512u) *reinterpret_cast<uint4*>(dst + k) = *reinterpret_cast<const uint4*>(st + k);
                if (skew) { const uint32_t k = skew + lane; if (k < 16u) dst[k] = st[k]; }
                { const uint32_t k = v1 + lane; if (k < end) dst[k] = st[k]; }
            } else {
                for (uint32_t k = skew + lane; k < end; k += 32u) dst[k] = st[k];
            }
            __syncwarp();                                                  // the stage is reused by the next range
        } else if (have) {
            k5_put_line(B, e, B.text + off);
        }
    }
    tr.end();
}

// ======================================================================================
// launchers (called by the host ABI layer)
// ======================================================================================
static int g_sm_count = 148;
static int g_k1_ctas_per_sm = 4;
static int g_k1a_ctas_per_sm = 8;          // resident CTAs of the screen kernel per SM (fewer leave room for the SA branch beside it)
void set_k1a_ctas_per_sm(int n) { g_k1a_ctas_per_sm = n < 1 ? 1 : (n > 8 ? 8 : n); }
static int g_k1_waves = 3;                 // grid = SMs x CTAs/SM x waves: > 1 trades prefetch depth for dynamic balance
void set_k1_waves(int n) { g_k1_waves = n < 1 ? 1 : (n > 16 ? 16 : n); }
void set_k1_ctas_per_sm(int n) { g_k1_ctas_per_sm = n < 1 ? 1 : (n > 4 ? 4 : n); }

size_t k1_flat_smem_bytes() { return sizeof(K1Smem); }

cudaError_t configure_kernels(int device)
{
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return e;
    g_sm_count = prop.multiProcessorCount;
    e = cudaFuncSetAttribute(k1_flat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Smem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k3b_sa_events, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K3bSmem));
}

// Launch with the programmatic-stream-serialization attribute: the kernel may be placed while its predecessor in the stream
// drains (it calls griddep_wait() before touching memory).  After anything but a kernel the attribute changes nothing.
template <class... KArgs, class... Args>
static void launch_dependent(void (*kernel)(KArgs...), uint32_t grid, uint32_t block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

void launch_k0(const DevBatch& B, const DevParams& P, cudaStream_t st)
{
    const uint32_t tiles = (B.n_reads + SCAN_TILE - 1) / SCAN_TILE;
    k0_classify<<<tiles, SCAN_THREADS, 0, st>>>(B, P);
}

// Raw-event layout for this submit: tile slices of 2^capt_log2 slots in the first half of the raw buffer at most,
// the rest is the atomically allocated overflow region.  Must be applied to the DevBatch before kernels 1 and 4b.
static constexpr uint32_t kRawSlabHeadroom = 256 * 1024;   // = kRawHeadroom in exlr_abi.cu

void plan_k1(DevBatch& B, int variant, uint32_t rpc, uint32_t* tiles_out)
{
    if (rpc < 1) rpc = 1;
    if (rpc > K1_MAX_RPC) rpc = K1_MAX_RPC;
    const uint32_t n_tiles = (B.n_reads + rpc - 1) / rpc;
    B.prim_slots = 0; B.capt_log2 = 0;
    B.slab = rpc >= 32 ? (uint32_t)K1_CAP : 16u;                         // overflow slab: small when tiles hold few records
    if (variant == 0) {
        int lg = 7;                                                       // up to K1_CAP = 128 slots per tile
        while (lg >= 0 && ((unsigned long long)n_tiles << lg) > (B.raw_cap - kRawSlabHeadroom) / 2) lg--;
        if (lg >= 0) { B.capt_log2 = (uint32_t)lg; B.prim_slots = n_tiles << lg; }
    }
    *tiles_out = n_tiles;
}

void launch_k1(const DevBatch& B, const DevParams& P, int variant, uint32_t rpc, cudaStream_t st)
{
    if (variant == 1) {
        const uint32_t blocks = min((B.n_reads + 7u) / 8u, (uint32_t)g_sm_count * 32u);
        k1_warp<<<blocks, 256, 0, st>>>(B, P);
    } else {
        if (rpc < 1) rpc = 1;
        if (rpc > K1_MAX_RPC) rpc = K1_MAX_RPC;
        const uint32_t n_tiles = (B.n_reads + rpc - 1) / rpc;
        uint32_t grid = min(n_tiles, (uint32_t)g_sm_count * (uint32_t)g_k1_ctas_per_sm * (uint32_t)g_k1_waves);
        if ((n_tiles + grid - 1) / grid > K1_MAX_TILES) grid = (n_tiles + K1_MAX_TILES - 1) / K1_MAX_TILES;
        k1_flat<<<grid, K1_THREADS, sizeof(K1Smem), st>>>(B, P, rpc, n_tiles, nullptr);
    }
}

// kernel 1c: the flat block scan over the long records kernel 1b listed, one record per tile.  The length of the list lives
// on the device; the grid covers the case that every record is on it.
void launch_k1c(const DevBatch& B, const DevParams& P, cudaStream_t st)
{
    // one resident wave (the list is usually short or empty, and every CTA of a larger grid costs a launch slot just to see that)
    uint32_t grid = min(B.n_reads, (uint32_t)g_sm_count * (uint32_t)g_k1_ctas_per_sm);
    if ((B.n_reads + grid - 1) / grid > K1_MAX_TILES) grid = (B.n_reads + K1_MAX_TILES - 1) / K1_MAX_TILES;
    launch_dependent(k1_flat, grid ? grid : 1u, K1_THREADS, sizeof(K1Smem), st, B, P, 1u, B.n_reads, (const uint32_t*)B.long_list);
}

// the screen pass (kernel 1a) and the resolution of its flagged steps (kernel 1b); n_ops < 2^32
uint32_t k1a_steps(unsigned long long n_ops) { return (uint32_t)((n_ops + K1A_STEP_OPS - 1) / K1A_STEP_OPS); }

void launch_k1a(const DevBatch& B, const DevParams& P, unsigned long long n_ops, cudaStream_t st)
{
    const uint32_t steps = k1a_steps(n_ops);
    uint32_t grid = (steps + K1A_THREADS / 32 - 1) / (K1A_THREADS / 32);
    const uint32_t cap = (uint32_t)g_sm_count * (uint32_t)g_k1a_ctas_per_sm;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    k1a_screen<<<grid, K1A_THREADS, 0, st>>>(B, P, n_ops);
}

void launch_k1b(const DevBatch& B, const DevParams& P, unsigned long long n_ops, bool use_k1c, cudaStream_t st)
{
    // the numbers of flagged steps and of claimed records live on the device: one resident wave strides over each list
    const uint32_t steps = k1a_steps(n_ops);
    const uint32_t g1 = min((steps + K1B_THREADS / 32 - 1) / (K1B_THREADS / 32), (uint32_t)g_sm_count * 8u);
    launch_dependent(k1b_claim, g1 ? g1 : 1u, K1B_THREADS, 0, st, B, P, n_ops, use_k1c ? 1u : 0u);
    const uint32_t g2 = min((B.n_reads + K1B_THREADS - 1) / K1B_THREADS, (uint32_t)g_sm_count * 8u);
    launch_dependent(k1b_walk, g2 ? g2 : 1u, K1B_THREADS, 0, st, B, P, use_k1c ? 1u : 0u);
}

void launch_k3a(const DevBatch& B, const DevParams& P, uint32_t mean_ops, cudaStream_t st)
{
    // grid-stride over a device-side count: size for the worst case, cap at one resident wave (spare CTAs cost launch time).
    // Lanes per record by the batch's mean CIGAR length: the more records a warp walks side by side, the fewer waves.
    const uint32_t cap = (uint32_t)g_sm_count * 8u;
    if (mean_ops <= 16) {
        const uint32_t g = min((B.n_reads + 127u) / 128u, cap);
        launch_dependent(k3a_sa_cigar<2>, g ? g : 1u, 256, 0, st, B, P);
    } else if (mean_ops <= 48) {
        const uint32_t g = min((B.n_reads + 63u) / 64u, cap);
        launch_dependent(k3a_sa_cigar<4>, g ? g : 1u, 256, 0, st, B, P);
    } else {
        const uint32_t g = min((B.n_reads + 31u) / 32u, cap);
        launch_dependent(k3a_sa_cigar<8>, g ? g : 1u, 256, 0, st, B, P);
    }
}

void launch_k3b(const DevBatch& B, const DevParams& P, cudaStream_t st)
{
    // tiles of 64 SA records, grid-stride; the SA-record count lives on the device, so the grid is sized from the batch but
    // capped at the CTAs that are resident at once: spare CTAs of an over-sized grid cost a launch slot each just to read the
    // count and leave, and a CTA with a second tile doubles the kernel's span
    const uint32_t gb = min((B.n_reads + K3B_TILE - 1u) / K3B_TILE, (uint32_t)g_sm_count * K3B_CTAS);
    launch_dependent(k3b_sa_events, gb ? gb : 1u, K3B_THREADS, sizeof(K3bSmem), st, B, P);
}

void launch_k4a(const DevBatch& B, const DevParams& P, cudaStream_t st)
{
    const uint32_t tiles = (B.n_reads + SCAN_TILE - 1) / SCAN_TILE;
    k4a_line_scan<<<tiles, SCAN_THREADS, 0, st>>>(B, P);
}

void launch_k4b(const DevBatch& B, const DevParams& P, cudaStream_t st)
{
    // the overflow count and the SA-record count live on the device: cover the largest they can be, capped at one resident wave
    uint32_t n = B.prim_slots > B.n_reads ? B.prim_slots : B.n_reads;
    const uint32_t room = B.raw_cap - B.prim_slots;
    if (room > n) n = room;
    const uint32_t grid = min((n + 255u) / 256u, (uint32_t)g_sm_count * 8u);
    launch_dependent(k4b_place, grid ? grid : 1u, 256, 0, st, B, P);
}

uint32_t scan_tiles(uint32_t n_reads) { return (n_reads + SCAN_TILE - 1) / SCAN_TILE; }
uint32_t text_scan_tiles(uint32_t max_events) { return (max_events + SCAN_THREADS - 1) / SCAN_THREADS + 1; }

void launch_k5(const DevBatch& B, cudaStream_t st)
{
    // the line count lives on the device: both kernels are one resident wave (5a draws its tiles from a ticket)
    const uint32_t ga = min((B.max_events + SCAN_THREADS - 1) / SCAN_THREADS, (uint32_t)g_sm_count * 8u);
    launch_dependent(k5a_line_bytes, ga ? ga : 1u, SCAN_THREADS, 0, st, B);
    const uint32_t gb = min((B.max_events + 255u) / 256u, (uint32_t)g_sm_count * 8u);
    launch_dependent(k5b_format, gb ? gb : 1u, 256, 0, st, B);
}

}  // namespace exlr
