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
    griddep_launch();
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
// result header: the 128-byte control block goes to (mapped, pinned) host memory by eight 16-byte stores of one warp, placed
// while the last kernel drains -- a cudaMemcpyAsync of 128 bytes costs ~9 us of stream time, this ~3
// ======================================================================================
__global__ void __launch_bounds__(32) k6_header(const Ctrl* ctrl, Ctrl* host_ctrl)
{
    griddep_wait();
    if (threadIdx.x < sizeof(Ctrl) / 16) reinterpret_cast<uint4*>(host_ctrl)[threadIdx.x] = reinterpret_cast<const uint4*>(ctrl)[threadIdx.x];
    __threadfence_system();
}

// ======================================================================================
// launchers
// ======================================================================================
static int g_sm_count = 148;
int sm_count() { return g_sm_count; }

cudaError_t configure_cigar_kernels();
cudaError_t configure_sa_kernels();

cudaError_t configure_kernels(int device)
{
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return e;
    g_sm_count = prop.multiProcessorCount;
    e = configure_cigar_kernels();
    if (e != cudaSuccess) return e;
    return configure_sa_kernels();
}

void launch_k0(const DevBatch& B, const DevParams& P, cudaStream_t st)
{
    const uint32_t tiles = (B.n_reads + SCAN_TILE - 1) / SCAN_TILE;
    k0_classify<<<tiles, SCAN_THREADS, 0, st>>>(B, P);
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
    const uint32_t grid = min((n + 255u) / 256u, (uint32_t)sm_count() * 8u);
    launch_dependent(k4b_place, grid ? grid : 1u, 256, 0, st, B, P);
}

void launch_header(const DevBatch& B, Ctrl* host_ctrl_dev, cudaStream_t st) { launch_dependent(k6_header, 1u, 32u, 0, st, (const Ctrl*)B.ctrl, host_ctrl_dev); }

uint32_t scan_tiles(uint32_t n_reads) { return (n_reads + SCAN_TILE - 1) / SCAN_TILE; }
uint32_t text_scan_tiles(uint32_t max_events) { return (max_events + SCAN_THREADS - 1) / SCAN_THREADS + 1; }

void launch_k5(const DevBatch& B, cudaStream_t st)
{
    // the line count lives on the device: both kernels are one resident wave (5a draws its tiles from a ticket)
    const uint32_t ga = min((B.max_events + SCAN_THREADS - 1) / SCAN_THREADS, (uint32_t)sm_count() * 8u);
    launch_dependent(k5a_line_bytes, ga ? ga : 1u, SCAN_THREADS, 0, st, B);
    const uint32_t gb = min((B.max_events + 255u) / 256u, (uint32_t)sm_count() * 8u);
    launch_dependent(k5b_format, gb ? gb : 1u, 256, 0, st, B);
}

}  // namespace exlr
