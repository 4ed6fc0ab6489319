// exlr_order.cu — classification and ordered output:
//
//   kernel 0  k0_classify    record filter (main.rs:169-190) + tid check (:198) + ordered list of
//                            kept records carrying an SA aux (:206), one chained scan
//   kernel 4a k4a_line_scan  pair-merge rule (main.rs:612-635) + far-edge domain check (:673-678)
//                            folded into the per-record line count, chained scan -> line offsets
//   kernel 4b k4b_place      ordered compaction: every event lands at its reference output
//                            position as a 48-byte exlr_event
//   kernel 5a k5a_line_bytes byte length of every output line -> chained scan -> byte offsets  (optional)
//   kernel 5b k5b_format     the lines themselves (utils.rs:225-236, 269-280)                  (optional)
#include "exlr_common.cuh"

namespace exlr {

// ======================================================================================
// kernel 0: filter + tid check + ordered SA-record list
// ======================================================================================
// WALK: batches of short CIGARs -- the thread that lists an SA record also walks its CIGAR (kernel 3a's work: 5 % of the records,
// ~30 ops each), which takes kernel 3a's launch off the SA chain.
template <bool WALK>
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
    const uint32_t at0 = at, listed = sa_mask;
    while (sa_mask) { const int i = __ffs(sa_mask) - 1; sa_mask &= sa_mask - 1; B.sa_list[at++] = r0 + i; }
    // kept counter: one atomic per warp
    for (int d = 16; d; d >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, d);
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(&B.ctrl->n_kept, kept);
    if (threadIdx.x == 0 && tile == (n + SCAN_TILE - 1) / SCAN_TILE - 1) B.ctrl->n_sa = grand;
    if (WALK) {
        uint32_t j = at0;
        for (uint32_t m = listed; m; m &= m - 1u) { SaSum sm; sa_cigar_sums_serial(B, r0 + (uint32_t)(__ffs((int)m) - 1), &sm); B.sa_sum[j++] = sm; }
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

// ---- the literal merge loop for more than two events (main.rs:636-742) ------------------------------------------------
// The loop is the identity unless some adjacent pair of Dels satisfies the far-edge predicate (main.rs:673-678), which needs
// merge_min > 2 * indel_min; kernel 1 flags such records (K1_FAR_HIT) and only those come here -- a rare path, run by the one
// thread that owns the record in kernel 4a (line count) and again in kernel 4b (the lines themselves).
// A round reads merge1 through a three-element window (prv, target, nxt) and appends to merge2, and the next round reads
// what this one appended: the (at most five) rounds are run as a pipeline of window stages fed one event at a time, so no
// event list is ever stored.  Pass 1 counts what every round emits; the lengths decide at which round the reference stops
// (`merge1.len() == merge2.len()` or five rounds) or panics (a round that starts with fewer than two events indexes out of
// bounds, main.rs:664-671); pass 2 re-runs the pipeline up to that round and writes its output.
struct MEv { uint32_t lend, rstart, rend, del; };                             // lstart = record.pos() for every event of a record
struct MStage { MEv a, b; uint32_t n_in, n_out; };
struct MEmit {                                                                // where the last round's events go (null: count only)
    exlr_event* dst; uint32_t n, cap, pos2, r, tid, neg;
    __device__ __forceinline__ void operator()(const MEv& e)
    {
        if (dst && n < cap)
            store_event(dst + n, (int64_t)pos2, (int64_t)e.lend, (int64_t)e.rstart, (int64_t)e.rend, r, tid, tid, EXLR_EV_META(1u, EXLR_KIND_INDEL, neg, neg));
        n++;
    }
};
static constexpr int kMergeRounds = 5;                                        // iter_times, main.rs:638

template <int K>
struct MFeed {
    // event e arrives as element n_in of round K+1's input; out-of-line so that the six call sites per round do not multiply
    static __device__ __noinline__ void push(MStage* st, uint32_t merge_min, int rounds, MEv e, MEmit* emit)
    {
        if (K >= rounds) { (*emit)(e); return; }
        MStage& s = st[K];
        s.n_in++;
        if (s.n_in >= 3u) {                                                   // idx = n_in - 2 (main.rs:662-671)
            const MEv P = s.a, T = s.b;
            const bool mp = P.del && T.del && abs_diff(P.lend, T.rstart) < merge_min;      // main.rs:673-675
            const bool mn = T.del && e.del && abs_diff(T.lend, e.rstart) < merge_min;      // main.rs:676-678
            if (mp || mn) {
                if (mp) { s.n_out++; MFeed<K + 1>::push(st, merge_min, rounds, MEv{P.lend, T.rstart, T.rend, 1u}, emit); }
                if (mn) { s.n_out++; MFeed<K + 1>::push(st, merge_min, rounds, MEv{T.lend, e.rstart, e.rend, 1u}, emit); }
            } else if (s.n_in == 3u) {                                        // idx == 1: prv, target, nxt (main.rs:724-727)
                s.n_out += 3u;
                MFeed<K + 1>::push(st, merge_min, rounds, P, emit);
                MFeed<K + 1>::push(st, merge_min, rounds, T, emit);
                MFeed<K + 1>::push(st, merge_min, rounds, e, emit);
            } else { s.n_out++; MFeed<K + 1>::push(st, merge_min, rounds, e, emit); }
        }
        s.a = s.b; s.b = e;
    }
};
template <>
struct MFeed<kMergeRounds> {
    static __device__ __forceinline__ void push(MStage*, uint32_t, int, MEv e, MEmit* emit) { (*emit)(e); }
};

// what the loop needs of one record, by value (a reference to the kernel's parameter block would make every thread of the
// kernel copy the block to its stack)
struct MArgs { const uint32_t* cigar; unsigned long long o0, o1; uint32_t pos2, total_consume, indel_min, merge_min; };

__device__ __forceinline__ MArgs merge_args(const DevBatch& B, const DevParams& P, uint32_t r, uint32_t total_consume)
{
    return MArgs{B.cigar, B.cigar_off[r], B.cigar_off[r + 1], (uint32_t)B.pos[r], total_consume, P.indel_min, P.merge_min};
}

// the record's indel events in CIGAR order (main.rs:549-600) through `rounds` rounds of the loop
static __device__ __noinline__ void merge_rounds(const MArgs A, int rounds, MStage* st, MEmit* emit)
{
    for (int k = 0; k < kMergeRounds; k++) { st[k].n_in = 0; st[k].n_out = 0; }
    uint32_t L = 0;
    for (unsigned long long o = A.o0; o < A.o1; o++) {
        const uint32_t v = __ldg(A.cigar + o), op = v & 15u, n = v >> 4;
        if (op == 2u) {
            if (n >= A.indel_min) MFeed<0>::push(st, A.merge_min, rounds, MEv{A.pos2 + L, A.pos2 + L + n, A.pos2 + A.total_consume, 1u}, emit);
            L += n;
        } else if (op == 1u) {
            if (n >= A.indel_min) MFeed<0>::push(st, A.merge_min, rounds, MEv{A.pos2 + L, A.pos2 + L, A.pos2 + L + n, 0u}, emit);
        } else if (op <= 8u && consumes_ref(op)) L += n;
    }
}

// Pass 1.  Returns lines | rounds << 28: the number of lines the loop leaves for the record and the round it stops at;
// 0xffffffff = the reference panics.
static __device__ __noinline__ uint32_t merge_plan(const MArgs A, uint32_t cnt)
{
    MStage st[kMergeRounds];
    MEmit none{nullptr, 0u, 0u, 0u, 0u, 0u, 0u};
    merge_rounds(A, kMergeRounds, st, &none);
    uint32_t in = cnt, lines = 0, rounds = 0;
    for (int k = 1; k <= kMergeRounds; k++) {
        if (in < 2u) return 0xffffffffu;                                      // merge1.len() - 2 wraps, merge1[idx] is out of bounds
        const uint32_t out = st[k - 1].n_out;                                 // (in == 2: no window, merge2 stays empty)
        lines = out; rounds = (uint32_t)k;
        if (in == out || k == kMergeRounds) break;                            // main.rs:733-737
        in = out;
    }
    return lines >= (1u << 28) ? 0xfffffffeu : (lines | (rounds << 28));
}

// Lines of one record without the >2 merge loop; *far is set when the loop has to decide (K1_FAR_HIT and more than two events).
__device__ __forceinline__ uint32_t record_lines(uint32_t csa, uint32_t info, bool* far)
{
    if (csa & CSA_DROP) return 0u;                                            // -k cap: no lines at all (main.rs:311-313)
    if (info & K1_MERGED) return (csa & CSA_CNT_MASK) + (info & K1_CNT_MASK);  // second pass: the loop's own line count
    if ((info & K1_CNT_MASK) > 2u && (info & K1_FAR_HIT)) *far = true;
    return (csa & CSA_CNT_MASK) + indel_lines(info);
}

// The rare path of kernel 4a, out of line (its call chain must not shape the register allocation of the scan around it):
// plans the literal merge loop for record r and leaves the number of indel lines it produces in the record's summary
// (K1_MERGED | lines: kernel 4a's second pass counts those, kernel 4b skips the record's raw events).
struct FarArgs { uint2* k1; uint2* far_list; Ctrl* ctrl; MArgs m; uint32_t r; };
static __device__ __noinline__ void far_record_plan(const FarArgs F)
{
    const uint2 k1 = F.k1[F.r];
    MArgs m = F.m; m.total_consume = k1.x;
    const uint32_t plan = merge_plan(m, k1.y & K1_CNT_MASK);
    uint32_t lines = 0;
    if (plan == 0xffffffffu) {
        // the reference panics here, after the SA arm of the same record has written its lines (main.rs:395-515 vs :664)
        report(F.ctrl, F.r, RANK_MERGE_DOMAIN);
    } else if (plan == 0xfffffffeu) F.ctrl->overflow = 1;                     // 2^28 lines from one record: no buffer holds that
    else {
        F.far_list[atomicAdd(&F.ctrl->n_far, 1u)] = make_uint2(F.r, plan >> 28);
        lines = plan & K1_CNT_MASK;
    }
    F.k1[F.r] = make_uint2(k1.x, K1_MERGED | lines);
}

// FAR = false is the kernel of every ordinary batch: the far-edge predicate cannot fire unless merge_min > 2 * indel_min (two
// Dels of at least indel_min lie between the edges it compares) or a coordinate wraps around 2^32, so the host launches this
// lean variant, whose registers are not shaped by the merge loop's call chain; should it meet such a record after all, it
// only raises need_far and the host runs the tail of the pipeline again with FAR = true.
template <bool FAR>
__global__ void __launch_bounds__(SCAN_THREADS, FAR ? 4 : 1) k4a_line_scan(DevBatch B, DevParams P)
{
    __shared__ uint32_t s_tile, s_warp[16];
    griddep_launch();                                  // kernel 4b may be placed; it waits for this grid before it reads anything
    CtaTrace tr(B, 5);
    if (threadIdx.x == 0) s_tile = atomicAdd(&B.ctrl->ticket_b, 1u);
    __syncthreads();
    const uint32_t tile = s_tile, n = B.n_reads;
    const uint32_t r0 = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t c[SCAN_ITEMS], mine, farmask;
    const bool full = r0 + SCAN_ITEMS <= n;
    for (;;) {
        mine = 0; farmask = 0;
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
                bool far = false;
                c[i] = record_lines(cs.u[i], info, &far); mine += c[i];
                farmask |= far ? 1u << i : 0u;
            }
        } else if (full) {
            union { uint4 v[4]; uint32_t u[16]; } cs;
            union { uint4 v[8]; uint2 p[16]; } k1;
#pragma unroll
            for (int k = 0; k < 4; k++) cs.v[k] = reinterpret_cast<const uint4*>(B.csa + r0)[k];
#pragma unroll
            for (int k = 0; k < 8; k++) k1.v[k] = reinterpret_cast<const uint4*>(B.k1 + r0)[k];
#pragma unroll
            for (int i = 0; i < SCAN_ITEMS; i++) {
                bool far = false;
                c[i] = record_lines(cs.u[i], k1.p[i].y, &far); mine += c[i];
                farmask |= far ? 1u << i : 0u;
            }
        } else {
#pragma unroll
            for (int i = 0; i < SCAN_ITEMS; i++) {
                const uint32_t r = r0 + i;
                uint32_t info = 0;
                if (r < n && (!B.k1_gated || ((B.dirty_bits[r >> 5] >> (r & 31u)) & 1u))) info = B.k1[r].y;
                bool far = false;
                c[i] = r < n ? record_lines(B.csa[r], info, &far) : 0u;
                mine += c[i];
                farmask |= far ? 1u << i : 0u;
            }
        }
        if (!farmask) break;
        if (!FAR) { B.ctrl->need_far = 1u; break; }
        // rare: records whose >2 merge loop is not the identity (main.rs:636-742).  The plan lands in the record's summary and
        // the counts are taken again (no dynamically indexed update of c[], which would move the array to local memory).
        if (FAR) {
            while (farmask) {
                const uint32_t r = r0 + (uint32_t)(__ffs((int)farmask) - 1);
                farmask &= farmask - 1u;
                far_record_plan(FarArgs{B.k1, B.far_list, B.ctrl, merge_args(B, P, r, 0u), r});
            }
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
// result header: the 128-byte control block goes to (mapped, pinned) host memory by eight 16-byte stores -- from the last CTA
// to finish of whichever kernel ends the step (4b, or 5b when the lines are formatted on the device).  A cudaMemcpyAsync of
// 128 bytes costs ~9 us of stream time, a kernel of its own ~4; this costs a ticket.  Call with the whole CTA, at its very end.
// ======================================================================================
__device__ __forceinline__ void finish_step(const DevBatch& B)
{
    __shared__ uint32_t s_last;
    __syncthreads();                                   // every thread of this CTA has issued its stores
    if (threadIdx.x == 0) { __threadfence(); s_last = atomicAdd(&B.ctrl->ticket_d, 1u) == gridDim.x - 1u; }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();                                   // ... and the other CTAs' are visible here
    Ctrl* ctrl = B.ctrl;
    if (threadIdx.x == 0) {
        const unsigned long long ek = __ldcg(&ctrl->err_key);
        if (ek) {
            // a panic in the indel arm comes after the SA arm's writes of the same record: kernel 4a counted exactly those lines for it
            const unsigned long long key = ~ek;
            const uint32_t r = (uint32_t)(key >> 8);
            ctrl->err_lines = (key & 0xffull) == RANK_MERGE_DOMAIN ? B.line_off[r + 1] - B.line_off[r] : 0u;
        }
        ctrl->ticket_d = 0;
        __threadfence();
    }
    __syncwarp();
    if (threadIdx.x < sizeof(Ctrl) / 16) reinterpret_cast<uint4*>(B.host_ctrl)[threadIdx.x] = __ldcg(reinterpret_cast<const uint4*>(ctrl) + threadIdx.x);
    __threadfence_system();
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
    if (k1.y & K1_MERGED) return;                                             // its lines come from the literal merge loop (far_list)
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
template <bool FAR>
__global__ void __launch_bounds__(256, 6) k4b_place(DevBatch B, DevParams P)
{
    griddep_wait();                                    // kernel 4a's line offsets
    griddep_launch();
    CtaTrace tr(B, 6);
    const uint32_t room = B.raw_cap - B.prim_slots, n_ovf = min(B.ctrl->n_raw, room), n_sa = B.ctrl->n_sa, n_far = FAR ? B.ctrl->n_far : 0u;
    tr.mid();
    const uint32_t limit = max(max(B.prim_slots, n_far), max(n_ovf, n_sa)), stride = gridDim.x * blockDim.x;
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
    // records the literal merge loop changed (rare): pass 2 writes the lines of the round kernel 4a found the loop stops at
    if constexpr (FAR) for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < n_far; x += stride) {
        const uint2 f = B.far_list[x];
        const uint32_t r = f.x, at = B.line_off[r] + (B.csa[r] & CSA_CNT_MASK);
        MStage st[kMergeRounds];
        MEmit emit{B.events + at, 0u, at < B.max_events ? B.max_events - at : 0u, (uint32_t)B.pos[r], r, (uint32_t)B.tid[r], (B.flag[r] >> 4) & 1u};
        merge_rounds(merge_args(B, P, r, B.k1[r].x), (int)f.y, st, &emit);
    }
    tr.end();
    if (!B.text_off) finish_step(B);                   // (with device formatting kernel 5b ends the step)
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

// the -v columns: "\t{tag}\t{qname}\tstrand:{record strand}\tflag:{flags}" (utils.rs:205-223, 252-267); the tag names the f.write site
__constant__ char kTagText[5][56] = {"excord-lr-alignment-event", "excord-lr-alignment-event-large-ins", "excord-lr-alignment-event-large-ins-one-alignments",
                                     "excord-lr-alignment-event-large-ins-two-alignments", "excord-lr-split-read"};
__constant__ uint32_t kTagLen[5] = {25, 35, 50, 50, 20};

__device__ __forceinline__ uint32_t verbose_len(const DevBatch& B, const exlr_event& e)
{
    const uint32_t r = e.read_idx, kind = min(EXLR_EV_KIND(e.meta), 4u), flag = B.flag[r];
    return 1u + kTagLen[kind] + 1u + (B.qname_off[r + 1] - B.qname_off[r]) + 8u + ((flag & 0x10u) ? 2u : 1u) + 6u + dec_len((int64_t)flag);
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
            if (B.verbose) len += verbose_len(B, e);
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
    p = put_dec(p, (int64_t)EXLR_EV_NUM(e.meta));
    if (B.verbose) {
        const uint32_t r = e.read_idx, kind = min(EXLR_EV_KIND(e.meta), 4u), flag = B.flag[r];
        *p++ = '\t';
        for (uint32_t k = 0; k < kTagLen[kind]; k++) *p++ = (uint8_t)kTagText[kind][k];
        *p++ = '\t';
        for (uint32_t k = B.qname_off[r]; k < B.qname_off[r + 1]; k++) *p++ = B.qnames[k];
        const char* s1 = "\tstrand:";
        for (int k = 0; k < 8; k++) *p++ = (uint8_t)s1[k];
        p = put_dec(p, (flag & 0x10u) ? -1 : 1);
        const char* s2 = "\tflag:";
        for (int k = 0; k < 6; k++) *p++ = (uint8_t)s2[k];
        p = put_dec(p, (int64_t)flag);
    }
    *p++ = '\n';
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
    griddep_launch();
    CtaTrace tr(B, 13);
    const uint32_t n = min(B.ctrl->n_events, B.max_events), lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    const bool fits = B.ctrl->text_bytes <= B.text_cap;  // else the host falls back to its own formatter (or grows the batch)
    for (uint32_t i0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32u; fits && i0 < n; i0 += nw * 32u) {
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
    finish_step(B);
}

// ======================================================================================
// launchers
// ======================================================================================

cudaError_t configure_cigar_kernels();
cudaError_t configure_sa_kernels();

cudaError_t configure_kernels(int device, int* sm_count_out)
{
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return e;
    *sm_count_out = prop.multiProcessorCount;
    e = configure_cigar_kernels();
    if (e != cudaSuccess) return e;
    return configure_sa_kernels();
}

void launch_k0(const DevBatch& B, const DevParams& P, bool walk, cudaStream_t st)
{
    const uint32_t tiles = (B.n_reads + SCAN_TILE - 1) / SCAN_TILE;
    if (walk) k0_classify<true><<<tiles, SCAN_THREADS, 0, st>>>(B, P);
    else k0_classify<false><<<tiles, SCAN_THREADS, 0, st>>>(B, P);
}

void launch_k4a(const DevBatch& B, const DevParams& P, bool far, cudaStream_t st)
{
    const uint32_t tiles = (B.n_reads + SCAN_TILE - 1) / SCAN_TILE;
    if (far) k4a_line_scan<true><<<tiles, SCAN_THREADS, 0, st>>>(B, P);
    else k4a_line_scan<false><<<tiles, SCAN_THREADS, 0, st>>>(B, P);
}

// before kernels 4a.. run a second time on the same batch (need_far): their tickets, scan status words and counters
__global__ void k_reset_tail(DevBatch B, uint32_t tiles_b, uint32_t tiles_c)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t k = i; k < tiles_b; k += stride) B.scan_b[k] = 0ull;
    for (uint32_t k = i; k < tiles_c; k += stride) B.scan_c[k] = 0ull;
    if (i == 0) { Ctrl* c = B.ctrl; c->ticket_b = 0; c->ticket_c = 0; c->n_events = 0; c->text_bytes = 0; c->n_far = 0; c->need_far = 0; c->err_lines = 0; }
}

void launch_reset_tail(const DevBatch& B, cudaStream_t st)
{
    k_reset_tail<<<64, 256, 0, st>>>(B, scan_tiles(B.n_reads), B.text_off ? text_scan_tiles(B.max_events) : 0u);
}

void launch_k4b(const DevBatch& B, const DevParams& P, bool far, cudaStream_t st)
{
    // the overflow count and the SA-record count live on the device: cover the largest they can be, capped at one resident wave
    uint32_t n = B.prim_slots > B.n_reads ? B.prim_slots : B.n_reads;
    const uint32_t room = B.raw_cap - B.prim_slots;
    if (room > n) n = room;
    const uint32_t per_sm = far ? EXLR_RESIDENT_PER_SM(k4b_place<true>, 256, 0) : EXLR_RESIDENT_PER_SM(k4b_place<false>, 256, 0);
    const uint32_t grid = min((n + 255u) / 256u, (uint32_t)B.hc.sms * per_sm);
    if (far) launch_dependent(k4b_place<true>, grid ? grid : 1u, 256, 0, st, B, P);
    else launch_dependent(k4b_place<false>, grid ? grid : 1u, 256, 0, st, B, P);
}


uint32_t scan_tiles(uint32_t n_reads) { return (n_reads + SCAN_TILE - 1) / SCAN_TILE; }
uint32_t text_scan_tiles(uint32_t max_events) { return (max_events + SCAN_THREADS - 1) / SCAN_THREADS + 1; }

void launch_k5(const DevBatch& B, cudaStream_t st)
{
    // the line count lives on the device: both kernels are one resident wave (5a draws its tiles from a ticket)
    const uint32_t ga = min((B.max_events + SCAN_THREADS - 1) / SCAN_THREADS, (uint32_t)B.hc.sms * EXLR_RESIDENT_PER_SM(k5a_line_bytes, SCAN_THREADS, 0));
    launch_dependent(k5a_line_bytes, ga ? ga : 1u, SCAN_THREADS, 0, st, B);
    const uint32_t gb = min((B.max_events + 255u) / 256u, (uint32_t)B.hc.sms * EXLR_RESIDENT_PER_SM(k5b_format, 256, 0));
    launch_dependent(k5b_format, gb ? gb : 1u, 256, 0, st, B);
}

}  // namespace exlr
