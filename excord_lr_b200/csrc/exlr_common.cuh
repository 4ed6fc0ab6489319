// exlr_common.cuh — device helpers shared by the kernel translation units (exlr_cigar.cu, exlr_sa.cu, exlr_order.cu):
// the record filter, the device error word, programmatic-dependent-launch fences, %globaltimer traces, and the chained
// (decoupled look-back) scan used by kernels 0, 4a and 5a.
#pragma once
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

// Kernel 3a's work for one record by one thread: per-op-type wrapping sums of the record's own CIGAR (main.rs:243-296), reference
// span D + M + = + X (split_read_event.rs:23-28), first-match offset (utils.rs:12-42).  For batches of short CIGARs, where kernel 0
// does this for the few SA records it lists instead of a launch of its own.
__device__ __forceinline__ void sa_cigar_sums_serial(const DevBatch& B, uint32_t r, SaSum* out)
{
    const unsigned long long o0 = B.cigar_off[r], o1 = B.cigar_off[r + 1];
    uint32_t sS = 0, sH = 0, sD = 0, sM = 0, sE = 0, sX = 0, bad = 0;
    unsigned long long ffm = 0; bool seenM = false;
    for (unsigned long long o = o0; o < o1; o += 8) {
        uint32_t vv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) vv[u] = o + u < o1 ? __ldg(B.cigar + o + u) : 0xfu;       // 0xf: not an op
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t v = vv[u], op = v & 15u, len = v >> 4;
            if (o + u < o1 && op > 8u) bad = 1;                                               // rust-htslib panics on an unknown op (main.rs:243)
            sM += op == 0u ? len : 0u; sD += op == 2u ? len : 0u; sS += op == 4u ? len : 0u;
            sH += op == 5u ? len : 0u; sE += op == 7u ? len : 0u; sX += op == 8u ? len : 0u;
            if (op == 0u) seenM = true;                                                       // utils.rs:28-29 stops at the first M
            if (!seenM && (op == 4u || op == 1u || op == 8u || op == 7u)) ffm += len;         // S I X =  (utils.rs:33)
        }
    }
    if (bad) report(B.ctrl, r, RANK_CIGAR_OP);
    out->S = sS; out->H = sH; out->refspan = (int64_t)sD + (int64_t)sM + (int64_t)sE + (int64_t)sX; out->ffm = (int64_t)ffm;
    out->pad[0] = out->pad[1] = 0;
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

// ---- host side: launch helpers shared by the launchers of every translation unit -------------------------------------

// CTAs of `kernel` that fit one SM at once (asked of the runtime once per call site).  Grids whose work count lives on the device
// are ONE resident wave striding over that count: a grid larger than what is resident runs a second wave whose CTAs only start
// when the first ones finish (measured on kernel 3b: 1.03 ms instead of 0.61 ms when its grid assumed 7 CTAs per SM and 6 fit).
#define EXLR_RESIDENT_PER_SM(kernel, threads, smem)                                                                             \
    ([&]() -> uint32_t {                                                                                                        \
        static int cached = 0;                                                                                                  \
        if (!cached) { int n = 0; if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, (int)(threads), (size_t)(smem)) != cudaSuccess || n < 1) n = 1; cached = n; } \
        return (uint32_t)cached;                                                                                                \
    }())

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

}  // namespace exlr
