// exlr_cigar.cu — the CIGAR path (reference src/main.rs:523-600 and the merge predicates of :612-635, :673-678):
//
//   kernel 1a k1a_screen     streams the CIGAR array once at HBM speed and lists the 512-op steps that
//                            hold an I/D >= indel_min (main.rs:553,569) or an unknown op code
//   kernel 1b k1b_claim      listed steps -> records (one owner per record), short / long lists
//             k1b_walk       one thread per short record, everything thread-local
//   kernel 1c k1_flat(list)  the long records, one per tile, by the flat block scan
//   kernel 1  k1_flat        flat TMA-staged block scan of everything (event-dense batches, EXLR_OPT_CIGAR_KERNEL=2):
//                            cp.async.bulk + mbarrier ring, one block prefix sum of the reference-consuming lengths,
//                            record boundaries resolved from the staged offsets
//             k1_warp        warp-per-record variant (A/B measurement; the odd long record of a short-record batch)
//
// HBM-bound integer work: no tensor cores anywhere on this path.
#include "exlr_common.cuh"

namespace exlr {

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
// SUMS: also the step's reference-consuming length (u32 wrapping, main.rs:528-545) -- what lets kernel 1d place an event of a
// long record without rescanning the record.
template <bool SUMS>
__device__ __forceinline__ bool k1a_step(uint32_t imin16, uint32_t lut_lo, uint32_t lut_hi, const uint4 (&q)[K1A_VEC], uint32_t* step_sum)
{
    uint32_t sus = 0, flags = 0, tsum = 0;
#pragma unroll
    for (int k = 0; k < K1A_VEC; k++) {
        const uint32_t vv[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t f;
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(f) : "r"(lut_lo), "r"(lut_hi), "r"(vv[j]));
            flags |= f;
            asm("{\n.reg .pred p;\nsetp.ge.u32 p, %1, %2;\n@p or.b32 %0, %0, %3;\n}" : "+r"(sus) : "r"(vv[j]), "r"(imin16), "r"(f));
            if (SUMS) asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(tsum) : "r"(f & 1u), "r"(vv[j] >> 4));
        }
    }
    if (SUMS) {
#pragma unroll
        for (int d = 16; d; d >>= 1) tsum += __shfl_xor_sync(0xffffffffu, tsum, d);
        *step_sum = tsum;
    }
    return __any_sync(0xffffffffu, ((sus & 2u) | (flags & 0x40u)) != 0u);
}

// DEEP: half the CTAs per SM, two warp steps' loads in flight per thread (the next step is requested before this one is screened):
// the same bytes in flight per SM from half the thread slots, which leaves the other half to whatever else is resident -- the
// SA branch beside it, or the kernels of other batches when several steps are in flight (EXLR_OPT_K1A_CTAS_PER_SM <= 4).
template <bool SUMS, bool DEEP>
__global__ void __launch_bounds__(K1A_THREADS, DEEP ? 4 : K1A_CTAS) k1a_screen(DevBatch B, DevParams P, unsigned long long n_ops)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nthreads = gridDim.x * K1A_THREADS, gt = blockIdx.x * K1A_THREADS + threadIdx.x;
    griddep_launch();                                  // kernel 1b may be placed; it waits for this grid before it reads anything
    CtaTrace tr(B, 8);
    const uint32_t imin16 = P.indel_min >= (1u << 28) ? 0xffffffffu : (P.indel_min << 4);
    const uint32_t nvec = (uint32_t)((n_ops + 3ull) / 4ull);               // < 2^31 (host); the cigar buffer is padded by 16 bytes
    const uint4* cig = reinterpret_cast<const uint4*>(B.cigar);
    // Warp steps are dealt round robin over all warps of the grid (event-dense genome regions are contiguous in the array).
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
    auto load = [&](uint4 (&q)[K1A_VEC], uint32_t v0) {
        const uint4* p = cig + v0 + lane;
        if ((unsigned long long)(v0 + 32 * K1A_VEC) * 4ull <= n_ops) {      // every op of the step exists
#pragma unroll
            for (int k = 0; k < K1A_VEC; k++) q[k] = __ldg(p + 32 * k);
        } else {                                                            // the last step: what lies past the last op reads as 0M
#pragma unroll
            for (int k = 0; k < K1A_VEC; k++) {
                const uint32_t vi = v0 + 32 * k + lane;
                q[k] = vi < nvec ? __ldg(p + 32 * k) : make_uint4(0u, 0u, 0u, 0u);
                const unsigned long long e0 = (unsigned long long)vi * 4ull;
                if (e0 + 1 > n_ops) q[k].x = 0u;
                if (e0 + 2 > n_ops) q[k].y = 0u;
                if (e0 + 3 > n_ops) q[k].z = 0u;
                if (e0 + 4 > n_ops) q[k].w = 0u;
            }
        }
    };
    uint32_t hitmask = 0, it = 0;
    auto screen = [&](const uint4 (&q)[K1A_VEC], uint32_t v0) {
        uint32_t ssum = 0;
        const bool hit = k1a_step<SUMS>(imin16, lut_lo, lut_hi, q, &ssum);
        if (SUMS && lane == 0) { const uint32_t st = v0 / (32 * K1A_VEC); B.step_sum[st] = ssum; B.step_flag[st] = hit ? 1 : 0; }
        hitmask |= (hit ? 1u : 0u) << (it & 31u);
        if ((it & 31u) == 31u) { append(hitmask, it - 31u); hitmask = 0; }
        it++;
    };
    if (!DEEP) {
        for (uint32_t v0 = gw * (32 * K1A_VEC); v0 < nvec; v0 += wstep) {
            uint4 q[K1A_VEC];
            load(q, v0);
            screen(q, v0);
        }
    } else {
        uint4 qa[K1A_VEC], qb[K1A_VEC];
        uint32_t v0 = gw * (32 * K1A_VEC);
        if (v0 < nvec) load(qa, v0);
        while (v0 < nvec) {                                                 // two steps per trip: each is screened while the other's loads fly
            const uint32_t v1 = v0 + wstep;                                 // (v0 + wstep cannot wrap: nvec < 2^31 and wstep < 2^26)
            if (v1 < nvec) load(qb, v1);
            screen(qa, v0);
            if (v1 >= nvec) break;
            const uint32_t v2 = v1 + wstep;
            if (v2 < nvec) load(qa, v2);
            screen(qb, v1);
            v0 = v2;
        }
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
static constexpr uint32_t K1B_LONG = 256;              // ops; longer records are scanned by kernel 1c / 1d or by a warp
static constexpr uint32_t K1B_LONG_BATCH_CUT = 64;     // the same cut in a batch of long records (kernels 1c / 1d take the rest)
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
            // by CIGAR length: a thread walks it (short_list); longer ones are scanned block-wide by kernel 1c (long_list) in
            // batches of long records, and by a warp each (warp_list) in batches of short records, where kernel 1c is not
            // launched.  (Measured on the ONT config: a warp per 3k-op record is latency-bound at ~1.8 us per 32 ops -- 1.1 ms for
            // the 58k records kernel 1c scans in 0.5 ms.)
            const unsigned long long len = o1 - o0;
            // (in a batch of long records the thread walk is kept for really short ones: a 256-op walk is 16 dependent load rounds)
            const uint32_t cut = use_k1c ? K1B_LONG_BATCH_CUT : K1B_LONG;
            const bool c_long = mine && len > cut && use_k1c, c_warp = mine && len > cut && !use_k1c, c_short = mine && len <= cut;
            const uint32_t sm = __ballot_sync(0xffffffffu, c_short), wm = __ballot_sync(0xffffffffu, c_warp), lm = __ballot_sync(0xffffffffu, c_long);
            uint32_t base = 0;
            if (lane == 0 && sm) base = atomicAdd(&B.ctrl->n_short, (uint32_t)__popc(sm));
            if (lane == 1 && wm) base = atomicAdd(&B.ctrl->n_warp, (uint32_t)__popc(wm));
            if (lane == 2 && lm) base = atomicAdd(&B.ctrl->n_long, (uint32_t)__popc(lm));
            const uint32_t sbase = __shfl_sync(0xffffffffu, base, 0), wbase = __shfl_sync(0xffffffffu, base, 1), lbase = __shfl_sync(0xffffffffu, base, 2);
            const uint32_t below = (1u << lane) - 1u;
            if (c_short) B.short_list[sbase + __popc(sm & below)] = r;
            if (c_warp) B.warp_list[wbase + __popc(wm & below)] = r;
            if (c_long) B.long_list[lbase + __popc(lm & below)] = r;
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

__global__ void __launch_bounds__(K1B_THREADS) k1b_walk(DevBatch B, DevParams P)
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
        bool is_long = false;                                                  // more events than parking space
        if (have) {
            r = B.short_list[i];
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
        if (have) {
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
        for (uint32_t lm = __ballot_sync(0xffffffffu, is_long); lm; lm &= lm - 1u)          // rare: the whole warp scans such a record
            k1_warp_record(B, P, __shfl_sync(0xffffffffu, r, __ffs((int)lm) - 1));
    }
    // the medium-length records: one warp each, 32 ops per step with a shuffle scan
    const uint32_t n_warp = B.ctrl->n_warp, nw = (gridDim.x * K1B_THREADS) >> 5;
    for (uint32_t i = (blockIdx.x * K1B_THREADS + t) >> 5; i < n_warp; i += nw) k1_warp_record(B, P, B.warp_list[i]);
    tr.end();
}

// ======================================================================================
// kernel 1d: the long records of a long-record batch, one warp each, WITHOUT rescanning them
//
// Kernel 1a<SUMS> left two things per 512-op step: its reference-consuming length and whether it holds an event candidate.  For a
// record that spans steps s0..s1 that is all that is needed: left_consume of an event = (partial sum of s0 inside the record) +
// (sums of the whole steps in between) + (prefix inside the event's own step); total_consume likewise.  So a warp scans only the
// record's two ragged end steps and its flagged steps (for an ONT record with one event: ~3 of its ~7..200 steps), 512 ops per
// scan with four 128-bit loads per lane, and events come out in CIGAR order with their sequence numbers and merge predicates
// (main.rs:523-600, 612-635, 673-678).  Measured on the ONT config: the flat block scan of the same records (kernel 1c) took 527 us.
// ======================================================================================
struct K1dState { uint32_t cnt, info, has_prev, pL, pn, pdel; };   // warp-uniform: events so far, flags, the previous event
// what a step scan needs of the batch, by value (the scan is out of line to bound the kernel's registers; a reference to the
// kernel's whole parameter block would make every thread copy it to its stack)
struct K1dArgs { const uint32_t* cigar; RawEv* raw; Ctrl* ctrl; uint32_t prim_slots, raw_cap, merge_min, imin16; };

// Scan the part of step `s` that lies inside the record's ops [o0, o1).  Lane l holds ops s*512 + 16 l .. + 15.
// Returns the reference-consuming length of that part.  EMIT: also writes the raw events of the part, in order, with
// `carry` = left_consume at the part's first op.
template <bool EMIT>
__device__ __noinline__ uint32_t k1d_scan_step(const K1dArgs A, uint32_t s, unsigned long long o0, unsigned long long o1, uint32_t carry,
                                               uint32_t pos2, uint32_t r, K1dState& S)
{
    const uint32_t lane = threadIdx.x & 31;
    const unsigned long long sb = (unsigned long long)s * K1A_STEP_OPS;                       // the step's first op
    // the record's part of the step as op numbers inside the step, [lo, hi) within 0..512
    const uint32_t lo = o0 > sb ? (uint32_t)(o0 - sb) : 0u, hi = o1 < sb + K1A_STEP_OPS ? (uint32_t)(o1 - sb) : K1A_STEP_OPS;
    const uint32_t e0 = lane * 16u;                                                             // my first op inside the step
    const uint4* c4 = reinterpret_cast<const uint4*>(A.cigar) + (sb >> 2) + lane * 4u;
    uint32_t v[16];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool any = e0 + 4u * k < hi && e0 + 4u * k + 4u > lo;                            // the vector overlaps the record
        const uint4 q = any ? __ldg(c4 + k) : make_uint4(0u, 0u, 0u, 0u);
        v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
    uint32_t tot = 0, evm = 0, flags = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        if (e0 + i - lo >= hi - lo) v[i] = 0u;                                                  // outside the record: reads as 0M
        uint32_t f;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(f) : "r"(K1_LUT_LO), "r"(K1_LUT_HI), "r"(v[i]));
        flags |= f;
        tot += (f & 1u) ? (v[i] >> 4) : 0u;
        if ((f & 2u) && v[i] >= A.imin16) evm |= 1u << i;
    }
    uint32_t incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
    const uint32_t part = __shfl_sync(0xffffffffu, incl, 31);
    if (!EMIT) return part;
    if (__any_sync(0xffffffffu, (flags & 0x40u) != 0u)) { if (flags & 0x40u) report(A.ctrl, r, RANK_CIGAR_OP); }   // rust-htslib panics on an unknown op
    const uint32_t em = __ballot_sync(0xffffffffu, evm != 0u);
    if (!em) return part;
    // ranks and raw slots
    const uint32_t mine = __popc(evm);
    uint32_t rincl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, rincl, d); if (lane >= (uint32_t)d) rincl += o; }
    const uint32_t n_ev = __shfl_sync(0xffffffffu, rincl, 31);
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&A.ctrl->n_raw, n_ev);
    base = A.prim_slots + __shfl_sync(0xffffffffu, base, 0) + rincl - mine;
    // first walk: my last event, for the lane above me; second walk: emit with the previous-event chain
    uint32_t L = carry + incl - tot, lL = 0, ln = 0, ldel = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t f;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(f) : "r"(K1_LUT_LO), "r"(K1_LUT_HI), "r"(v[i]));
        if (evm & (1u << i)) { lL = L; ln = v[i] >> 4; ldel = (v[i] & 15u) == 2u; }
        L += (f & 1u) ? (v[i] >> 4) : 0u;
    }
    const uint32_t lower = em & ((1u << lane) - 1u);
    const int src = lower ? 31 - __clz(lower) : 0;
    uint32_t pL = __shfl_sync(0xffffffffu, lL, src), pn = __shfl_sync(0xffffffffu, ln, src), pdel = __shfl_sync(0xffffffffu, ldel, src);
    bool has_prev = lower != 0u;
    if (!has_prev && S.has_prev) { pL = S.pL; pn = S.pn; pdel = S.pdel; has_prev = true; }
    uint32_t seq = S.cnt + rincl - mine, info = 0, slot = base;
    L = carry + incl - tot;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint32_t f;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(f) : "r"(K1_LUT_LO), "r"(K1_LUT_HI), "r"(v[i]));
        if (evm & (1u << i)) {
            const uint32_t len = v[i] >> 4, del = (v[i] & 15u) == 2u;
            if (has_prev && del && pdel) {
                if (seq == 1u && abs_diff(pos2 + L, pos2 + pL + pn) < A.merge_min) info |= K1_PAIR_MERGE;   // main.rs:615
                if (abs_diff(pos2 + pL, pos2 + L + len) < A.merge_min) info |= K1_FAR_HIT;                  // main.rs:673-678
            }
            if (slot < A.raw_cap) {
                uint4* d = reinterpret_cast<uint4*>(A.raw + slot);
                d[0] = make_uint4(r, seq, L, len | (del << 31));
                d[1] = make_uint4(has_prev ? pL : 0u, 0u, 0u, 0u);
            } else A.ctrl->overflow = 1;
            pL = L; pn = len; pdel = del; has_prev = true; seq++; slot++;
        }
        L += (f & 1u) ? (v[i] >> 4) : 0u;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) info |= __shfl_xor_sync(0xffffffffu, info, d);
    const int top = 31 - __clz(em);                                                             // the step's last event becomes "previous"
    S.pL = __shfl_sync(0xffffffffu, lL, top); S.pn = __shfl_sync(0xffffffffu, ln, top); S.pdel = __shfl_sync(0xffffffffu, ldel, top);
    S.has_prev = 1; S.cnt += n_ev; S.info |= info;
    return part;
}

__global__ void __launch_bounds__(256, 3) k1d_long(DevBatch B, DevParams P)
{
    const uint32_t lane = threadIdx.x & 31;
    griddep_wait();                                    // k1b's long list
    griddep_launch();
    CtaTrace tr(B, 11);
    const uint32_t n_list = B.ctrl->n_long, nw = (gridDim.x * blockDim.x) >> 5;
    const K1dArgs A{B.cigar, B.raw, B.ctrl, B.prim_slots, B.raw_cap, P.merge_min, P.indel_min >= (1u << 28) ? 0xffffffffu : (P.indel_min << 4)};
    for (uint32_t li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; li < n_list; li += nw) {
        const uint32_t r = B.long_list[li];
        const unsigned long long o0 = B.cigar_off[r], o1 = B.cigar_off[r + 1];
        const uint32_t pos2 = (uint32_t)B.pos[r];
        const uint32_t s0 = (uint32_t)(o0 / K1A_STEP_OPS), s1 = (uint32_t)((o1 - 1ull) / K1A_STEP_OPS);       // o1 > o0: long records are not empty
        K1dState S{0u, 0u, 0u, 0u, 0u, 0u};
        uint32_t T;
        if (s0 == s1) T = k1d_scan_step<true>(A, s0, o0, o1, 0u, pos2, r, S);
        else {
            const uint32_t first = k1d_scan_step<false>(A, s0, o0, o1, 0u, pos2, r, S);
            const uint32_t last = k1d_scan_step<false>(A, s1, o0, o1, 0u, pos2, r, S);
            uint32_t carry = 0;
            for (uint32_t w0 = s0; w0 <= s1; w0 += 32u) {                                        // the record's steps, 32 at a time
                const uint32_t s = w0 + lane;
                const bool in = s <= s1;
                const uint32_t val = !in ? 0u : (s == s0 ? first : (s == s1 ? last : B.step_sum[s]));
                const bool flagged = in && B.step_flag[s] != 0;
                uint32_t incl = val;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= (uint32_t)d) incl += o; }
                for (uint32_t m = __ballot_sync(0xffffffffu, flagged); m; m &= m - 1u) {            // only the flagged steps are looked at
                    const int b = __ffs((int)m) - 1;
                    const uint32_t at = carry + __shfl_sync(0xffffffffu, incl - val, b);
                    k1d_scan_step<true>(A, w0 + (uint32_t)b, o0, o1, at, pos2, r, S);
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
                if (w0 + 32u < w0) break;                                                        // (u32 wrap guard)
            }
            T = carry;
        }
        if (lane == 0) B.k1[r] = make_uint2(T, (S.cnt & K1_CNT_MASK) | S.info);
    }
    tr.end();
}

// ======================================================================================
// launchers
// ======================================================================================

size_t k1_flat_smem_bytes() { return sizeof(K1Smem); }
cudaError_t configure_cigar_kernels() { return cudaFuncSetAttribute(k1_flat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Smem)); }

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
        const uint32_t blocks = min((B.n_reads + 7u) / 8u, (uint32_t)B.hc.sms * 32u);
        k1_warp<<<blocks, 256, 0, st>>>(B, P);
    } else {
        if (rpc < 1) rpc = 1;
        if (rpc > K1_MAX_RPC) rpc = K1_MAX_RPC;
        const uint32_t n_tiles = (B.n_reads + rpc - 1) / rpc;
        uint32_t grid = min(n_tiles, (uint32_t)B.hc.sms * (uint32_t)B.hc.k1_ctas * (uint32_t)B.hc.k1_waves);
        if ((n_tiles + grid - 1) / grid > K1_MAX_TILES) grid = (n_tiles + K1_MAX_TILES - 1) / K1_MAX_TILES;
        k1_flat<<<grid, K1_THREADS, sizeof(K1Smem), st>>>(B, P, rpc, n_tiles, nullptr);
    }
}

// kernel 1c: the flat block scan over the long records kernel 1b listed, one record per tile.  The length of the list lives
// on the device; the grid covers the case that every record is on it.
void launch_k1c(const DevBatch& B, const DevParams& P, cudaStream_t st)
{
    // one resident wave (the list is usually short or empty, and every CTA of a larger grid costs a launch slot just to see that)
    uint32_t grid = min(B.n_reads, (uint32_t)B.hc.sms * (uint32_t)B.hc.k1_ctas);
    if ((B.n_reads + grid - 1) / grid > K1_MAX_TILES) grid = (B.n_reads + K1_MAX_TILES - 1) / K1_MAX_TILES;
    launch_dependent(k1_flat, grid ? grid : 1u, K1_THREADS, sizeof(K1Smem), st, B, P, 1u, B.n_reads, (const uint32_t*)B.long_list);
}

// kernel 1d: one warp per listed long record, one resident wave
void launch_k1d(const DevBatch& B, const DevParams& P, cudaStream_t st)
{
    const uint32_t grid = min((B.n_reads + 7u) / 8u, (uint32_t)B.hc.sms * EXLR_RESIDENT_PER_SM(k1d_long, 256, 0));
    launch_dependent(k1d_long, grid ? grid : 1u, 256u, 0, st, B, P);
}

// the screen pass (kernel 1a) and the resolution of its flagged steps (kernel 1b); n_ops < 2^33
uint32_t k1a_steps(unsigned long long n_ops) { return (uint32_t)((n_ops + K1A_STEP_OPS - 1) / K1A_STEP_OPS); }

void launch_k1a(const DevBatch& B, const DevParams& P, unsigned long long n_ops, bool sums, cudaStream_t st)
{
    const uint32_t steps = k1a_steps(n_ops);
    uint32_t grid = (steps + K1A_THREADS / 32 - 1) / (K1A_THREADS / 32);
    const uint32_t cap = (uint32_t)B.hc.sms * (uint32_t)B.hc.k1a_ctas;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    const bool deep = B.hc.k1a_ctas <= 4;              // half the thread slots, two steps' loads in flight per thread
    if (sums) { if (deep) k1a_screen<true, true><<<grid, K1A_THREADS, 0, st>>>(B, P, n_ops); else k1a_screen<true, false><<<grid, K1A_THREADS, 0, st>>>(B, P, n_ops); }
    else { if (deep) k1a_screen<false, true><<<grid, K1A_THREADS, 0, st>>>(B, P, n_ops); else k1a_screen<false, false><<<grid, K1A_THREADS, 0, st>>>(B, P, n_ops); }
}

void launch_k1b(const DevBatch& B, const DevParams& P, unsigned long long n_ops, bool use_k1c, cudaStream_t st)
{
    // the numbers of flagged steps and of claimed records live on the device: one resident wave strides over each list
    const uint32_t steps = k1a_steps(n_ops);
    const uint32_t g1 = min((steps + K1B_THREADS / 32 - 1) / (K1B_THREADS / 32), (uint32_t)B.hc.sms * EXLR_RESIDENT_PER_SM(k1b_claim, K1B_THREADS, 0));
    launch_dependent(k1b_claim, g1 ? g1 : 1u, K1B_THREADS, 0, st, B, P, n_ops, use_k1c ? 1u : 0u);
    const uint32_t g2 = min((B.n_reads + K1B_THREADS - 1) / K1B_THREADS, (uint32_t)B.hc.sms * EXLR_RESIDENT_PER_SM(k1b_walk, K1B_THREADS, 0));
    launch_dependent(k1b_walk, g2 ? g2 : 1u, K1B_THREADS, 0, st, B, P);
}

}  // namespace exlr
