// exlr_abi.cu — host side of libexlr_cuda.so: the extern "C" entry points of include/exlr.h.
//
// One exlr_ctx per GPU; each exlr_batch owns a CUDA stream, pinned host staging for the
// structure-of-arrays views, the device copies, scratch and result buffers, so several
// batches can be in flight (H2D of one overlapping the kernels / D2H of another).
// There is no CPU fallback anywhere in this file: without a usable sm_100 device every
// compute entry point returns EXLR_ERR_CUDA.
#include <algorithm>
#include <cstdint>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <sys/mman.h>
#include <cuda_runtime.h>

#include "exlr_device.cuh"

using namespace exlr;

static thread_local char g_cuda_err[512] = "";

static int cuda_fail(cudaError_t e, const char* what)
{
    snprintf(g_cuda_err, sizeof g_cuda_err, "%s: %s", what, cudaGetErrorString(e));
    return EXLR_ERR_CUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

// EXLR_ALLOC_TRACE=1 in the environment: every buffer allocation reports its size and duration on stderr (setup-time profiling)
static bool alloc_trace() { static const bool on = getenv("EXLR_ALLOC_TRACE") != nullptr; return on; }
template <class F>
static cudaError_t traced(const char* what, size_t bytes, F f)
{
    if (!alloc_trace()) return f();
    const auto t0 = std::chrono::steady_clock::now();
    const cudaError_t e = f();
    fprintf(stderr, "exlr alloc: %-28s %10.1f MB %8.2f ms\n", what, bytes / 1e6, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return e;
}
static cudaError_t dev_alloc(void** p, size_t bytes, const char* what) { return traced(what, bytes, [&] { return cudaMalloc(p, bytes); }); }
static cudaError_t host_alloc(void** p, size_t bytes, unsigned flags, const char* what) { return traced(what, bytes, [&] { return cudaHostAlloc(p, bytes, flags); }); }

// A large pinned buffer.  cudaHostAlloc pins 4 KB pages one by one inside the driver (~0.45 ms per MB measured, and kernel launches
// of other threads wait behind it); an anonymous mapping backed by transparent huge pages, faulted in here and then registered,
// costs a third of that and only the short cudaHostRegister runs inside the driver (tools/exp/pin_bench.cu: 128 MB in 22 ms instead
// of 57 ms, same 55 GB/s H2D).  Falls back to cudaHostAlloc wherever any step of that is refused.
struct Pinned { void* p = nullptr; void* map = nullptr; size_t map_bytes = 0; };
static cudaError_t pinned_alloc(Pinned* out, size_t bytes, const char* what)
{
    *out = Pinned{};
    return traced(what, bytes, [&]() -> cudaError_t {
        constexpr size_t kHuge = 2u << 20;
        static const bool no_thp = getenv("EXLR_NO_THP") != nullptr;
        if (bytes >= (8u << 20) && !no_thp) {
            const size_t body = (bytes + kHuge - 1) / kHuge * kHuge, len = body + kHuge;
            void* q = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (q != MAP_FAILED) {
                char* a = (char*)(((uintptr_t)q + kHuge - 1) / kHuge * kHuge);
                madvise(a, body, MADV_HUGEPAGE);
                for (size_t o = 0; o < bytes; o += 4096) ((volatile char*)a)[o] = 0;      // the page faults happen here, outside the driver
                if (cudaHostRegister(a, bytes, cudaHostRegisterDefault) == cudaSuccess) { out->p = a; out->map = q; out->map_bytes = len; return cudaSuccess; }
                cudaGetLastError();
                munmap(q, len);
            }
        }
        return cudaHostAlloc(&out->p, bytes, cudaHostAllocDefault);
    });
}
static void pinned_free(Pinned* b)
{
    if (b->map) { cudaHostUnregister(b->p); munmap(b->map, b->map_bytes); }
    else if (b->p) cudaFreeHost(b->p);
    *b = Pinned{};
}

enum { EV_START = 0, EV_H2D, EV_K0, EV_K1, EV_K3A, EV_K3B, EV_K4A, EV_K4B, EV_D2H, EV_COUNT };

struct exlr_ctx {
    int device = 0;
    exlr_params params{};
    DevParams dparams{};
    std::vector<std::string> ref_stripped;     // "chr" removed (aligments_event.rs:38-42)
    uint8_t* d_ref_bytes = nullptr; uint32_t* d_ref_off = nullptr; int n_ref = 0;
    int cigar_kernel = 0;                      // EXLR_OPT_CIGAR_KERNEL (0 auto, 1 warp, 2 flat scan, 3 screen + thread per record)
    int overlap = 1;                           // EXLR_OPT_OVERLAP: kernel 1 on a second stream beside the SA branch
    int trace = 0;                             // EXLR_OPT_TRACE: kernel 1 writes a per-CTA timeline (debug)
    int stage_timing = 1;                      // EXLR_OPT_STAGE_TIMING: CUDA events between the kernels (exlr_timing per stage)
    int k1_ctas = 0;                           // EXLR_OPT_K1_CTAS_PER_SM (0 = 3 when overlapping, which leaves room for the SA branch, else 4)
    int k1a_ctas = 8, k1_waves = 3, sms = 148; // EXLR_OPT_K1A_CTAS_PER_SM, EXLR_OPT_K1_WAVES; multiprocessors of `device`
    uint32_t reads_per_cta = 0;                // EXLR_OPT_READS_PER_CTA (0 = auto)
    int device_format = 0;                     // EXLR_OPT_DEVICE_FORMAT: kernels 5a/5b write the output lines; read with exlr_wait_text
    int verbose_text = 0;                      // EXLR_OPT_VERBOSE_TEXT: BAM batches format the -v columns too (the read names are on the device there)
    int long_records = 0;                      // EXLR_OPT_LONG_RECORDS: 0 auto (by mean CIGAR length), 1 never, 2 kernel 1c, 3 kernel 1d
    std::atomic<int> skip_screen{0};           // auto mode: batches left to run without the screen pass (the last screened one was event-dense);
                                               // written by whoever waits a batch, read by whoever submits the next (two threads in the CLI)
    std::atomic<uint64_t> ev_hint{0}, text_hint{0};   // events / text bytes of the last waited batch: how much exlr_submit copies back speculatively
    int bgzf_crc = 1;                          // EXLR_OPT_BGZF_CRC: kb_inflate verifies the CRC-32 of every BGZF block
    int k0_walk = 0;                           // EXLR_OPT_K0_WALK: kernel 0 does kernel 3a's work in batches of short CIGARs (measured slower: off)
    int wc_input = 0;                          // EXLR_OPT_WC_INPUT: pinned input views allocated write-combined
    int graph = 1;                             // EXLR_OPT_GRAPH: repeated shapes run as one CUDA graph launch
    int k3_fold = 0;                           // EXLR_OPT_K3_FOLD: 1 = kernel 3b does kernel 3a's work itself in batches of short CIGARs (measured slower: off)
    cudaEvent_t ev_origin = nullptr;           // recorded at exlr_create: the context's clock for exlr_bam_info.t_ms
    bool far_mode = false;                     // merge_min > 2 * indel_min: the >2 merge loop (main.rs:636-742) can change the events, kernels 4a/4b run their FAR variants
};

// What a submit is going to launch: decided on the host before anything is enqueued (and part of the identity of a captured graph).
struct StepPlan {
    bool overlap, fork_late, screened, long_batch, two_level, fold, k0_walk, far, formatted;
    int variant, trace; uint32_t rpc;
    unsigned long long n_reads, n_ops;
    const void* events;                        // (exlr_batch_grow moves the event buffers: a graph captured before it is stale)
    bool operator==(const StepPlan& o) const
    {
        return overlap == o.overlap && fork_late == o.fork_late && screened == o.screened && long_batch == o.long_batch && two_level == o.two_level && fold == o.fold && k0_walk == o.k0_walk &&
               far == o.far && formatted == o.formatted && variant == o.variant && trace == o.trace && rpc == o.rpc && n_reads == o.n_reads && n_ops == o.n_ops && events == o.events;
    }
};

struct exlr_batch {
    exlr_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: kernel 1 runs beside the SA branch (kernels 0, 3a, 3b)
    cudaEvent_t ev[EV_COUNT] = {};
    cudaEvent_t ev_fork = nullptr, ev_k1_begin = nullptr, ev_k1_mid = nullptr, ev_k1_end = nullptr;
    bool screened = false;                     // the last submit ran kernels 1a + 1b instead of kernel 1
    exlr_batch_views hv{};                     // pinned host views
    void* h_slab = nullptr;                    // pinned: inputs
    void* h_out = nullptr;                     // pinned: ctrl + line_off
    void* d_evslab = nullptr;                  // device: everything sized by max_events (raw, sa_ev, events, text); exlr_batch_grow replaces it
    uint64_t d2h_events = 0, d2h_text = 0;     // events / text bytes the last submit copied back behind its kernels (speculative, see run_kernels)
    bool have_line_off = false;                // ... and the line offsets
    uint64_t h2d_bytes = 0, d2h_bytes = 0;     // bytes the last submit + wait moved over PCIe
    Ctrl* h_ctrl = nullptr; uint32_t* h_line_off = nullptr; exlr_event* h_events = nullptr;
    Ctrl* h_ctrl_dev = nullptr;                // device address of h_ctrl (mapped pinned memory): the result header is stored there by a kernel
    char* h_text = nullptr;                    // pinned: formatted lines (allocated with the batch when EXLR_OPT_DEVICE_FORMAT is set)
    uint64_t h_text_cap = 0;                   // its size: dv.text_cap, or less for a lean batch (grown on demand by exlr_wait_text)
    bool lean_host = false;                    // a BAM batch: its results normally leave as formatted text, so the pinned copies of the
                                               // events and line offsets (sized by the worst-case record count) are only allocated when
                                               // exlr_wait is actually called, and the pinned text buffer starts small
    uint32_t* h_loff_full = nullptr;           // lean batch: the line offsets, once exlr_wait has asked for them
    Pinned pin_slab, pin_events, pin_text, pin_loff, pin_comp;   // how h_slab, h_events, h_text, h_loff_full, h_comp were pinned
    bool formatted = false;                    // the last submit ran kernels 5a/5b
    bool device_format = false;                // EXLR_OPT_DEVICE_FORMAT was set when the batch was allocated
    bool verbose_text = false;                 // a BAM batch allocated with EXLR_OPT_VERBOSE_TEXT: its device-formatted lines carry the -v columns
    void* d_slab = nullptr;                    // one device allocation, carved up below
    DevBatch dv{};
    size_t ctrl_bytes = 0;                     // ctrl + both scan status arrays (one memset)
    size_t scan_c_bytes = 0;                   // kernel 5a's scan status words (second memset, only with EXLR_OPT_DEVICE_FORMAT)
    uint64_t n_reads = 0, n_ops = 0;
    bool submitted = false, resident_uploaded = false, have_timing = false, stage_timed = false;
    bool far_ran = false;                      // the last submit ran the FAR variants of kernels 4a/4b
    cudaGraphExec_t gexec = nullptr;           // the step as a CUDA graph, once the same shape has been submitted twice in a row
    StepPlan gplan{}, last_plan{}; bool have_last_plan = false; uint32_t glaunches = 0;
    uint32_t launches = 0;
    unsigned long long* d_dbg = nullptr;
    // ---- BAM input decoded on the device (exlr_bam_*): compressed chunk + block table in pinned memory, the rest on the device
    bool is_bam = false;
    uint8_t* h_comp = nullptr; exlr_bgzf_block* h_blocks = nullptr; BgzfBlock* h_btab = nullptr; BamCtrl* h_bctrl = nullptr;
    void* d_bam = nullptr; DevBam db{}; BgzfBlock* d_btab = nullptr;
    uint64_t front_u = 0, bam_u_end = 0, bam_origin = 0, bam_info_tail = 0; uint32_t bam_n_new = 0, bam_n_front = 0;
    cudaEvent_t ev_tail_copied = nullptr;      // recorded on this batch's stream once it has copied the previous chunk's tail out of that batch's U
    cudaEvent_t tail_reader = nullptr;         // ... and, in that previous batch: the event its next submit has to wait for (it may live on another GPU)
    size_t bam_zero_bytes = 0;                 // BamCtrl + the three scan status arrays (one memset per submit)
    uint64_t max_comp = 0, u_cap = 0; uint32_t max_blocks = 0;
    int bam_state = 0;                         // 0 idle, 1 exlr_bam_submit done, 2 exlr_bam_walk done, 3 exlr_bam_extract done
    uint64_t bam_comp_bytes = 0;
    cudaEvent_t ev_bam[4] = {};                // start, after H2D, after inflate, after walk + gather
};

// kernel 1 reserves overflow slabs of 128 raw slots ahead of use (one spare per persistent CTA, see k1_flush): room for
// them on top of the caller's max_events so that only real events can exhaust the buffer
static constexpr size_t kRawHeadroom = 256 * 1024;
// EXLR_OPT_DEVICE_FORMAT: room for the formatted lines, per max_events entry (a typical line is 40-60 bytes; when a batch needs
// more, exlr_wait_text says so and the caller formats on the host)
static constexpr size_t kTextBytesPerLine = 96;
static constexpr size_t kTextBytesPerVerboseLine = 288;    // -v lines carry a tag of up to 50 bytes and the read name (BAM batches)

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static void drop_graph(exlr_batch* b)
{
    if (b->gexec) { cudaGraphExecDestroy(b->gexec); b->gexec = nullptr; }
    b->have_last_plan = false;
}


// Everything whose size follows max_events: the raw / SA / final event buffers, the text buffers of kernels 5a/5b with their
// scan status words, and the pinned host copies.  Separate from the input + per-record slab so that exlr_batch_grow can
// replace it while the packed records stay where they are.
static constexpr size_t kLeanTextBytes = 16u << 20;     // pinned text buffer a lean batch starts with

static void free_event_buffers(exlr_batch* b)
{
    cudaFree(b->d_evslab); b->d_evslab = nullptr;
    pinned_free(&b->pin_events); b->h_events = nullptr;
    pinned_free(&b->pin_text); b->h_text = nullptr;
}

static int alloc_event_buffers(exlr_batch* b, uint64_t max_events)
{
    if (max_events == 0 || max_events >= 0xfff00000ull) return EXLR_ERR_ARG;
    const bool fmt = b->device_format;
    const size_t text_cap = fmt ? max_events * (b->verbose_text ? kTextBytesPerVerboseLine : kTextBytesPerLine) : 0;
    if (text_cap >= 0xfffffff0ull) return EXLR_ERR_ARG;
    const uint32_t ttiles = fmt ? text_scan_tiles((uint32_t)max_events) : 0;
    size_t dof = 0;
    auto dcarve = [&](size_t bytes) { size_t at = dof; dof = align_up(dof + bytes, 256); return at; };
    const size_t d_scan = dcarve((size_t)ttiles * 8), d_raw = dcarve((max_events + kRawHeadroom) * sizeof(RawEv)),
                 d_saev = dcarve(max_events * sizeof(exlr_event)), d_toff = dcarve(fmt ? (max_events + 1) * 4 : 0),
                 d_text = dcarve(text_cap + 16), d_ev = dcarve(max_events * sizeof(exlr_event));
    cudaError_t e = dev_alloc(&b->d_evslab, dof, "device events+text");
    if (e == cudaSuccess && !b->lean_host) { e = pinned_alloc(&b->pin_events, max_events * sizeof(exlr_event), "pinned events"); b->h_events = (exlr_event*)b->pin_events.p; }
    b->h_text_cap = b->lean_host ? std::min<size_t>(text_cap, kLeanTextBytes) : text_cap;
    if (e == cudaSuccess && fmt) { e = pinned_alloc(&b->pin_text, b->h_text_cap + 16, "pinned text"); b->h_text = (char*)b->pin_text.p; }
    if (e != cudaSuccess) { free_event_buffers(b); return cuda_fail(e, "event buffers"); }
    char* ds = (char*)b->d_evslab;
    DevBatch& v = b->dv;
    v.scan_c = (unsigned long long*)(ds + d_scan); b->scan_c_bytes = (size_t)ttiles * 8;
    v.raw = (RawEv*)(ds + d_raw); v.raw_cap = (uint32_t)(max_events + kRawHeadroom);
    v.sa_ev = (exlr_event*)(ds + d_saev);
    v.text_off = fmt ? (uint32_t*)(ds + d_toff) : nullptr; v.text = (uint8_t*)(ds + d_text); v.text_cap = (uint32_t)text_cap;
    v.events = (exlr_event*)(ds + d_ev);
    v.max_events = (uint32_t)max_events;
    b->hv.max_events = max_events;
    return EXLR_OK;
}

// exlr_wait on a lean batch: the pinned copies it hands out are allocated now
static int ensure_fetch_buffers(exlr_batch* b)
{
    if (!b->lean_host) return EXLR_OK;
    cudaError_t e = cudaSuccess;
    if (!b->h_events) { e = pinned_alloc(&b->pin_events, b->hv.max_events * sizeof(exlr_event), "pinned events (on demand)"); b->h_events = (exlr_event*)b->pin_events.p; }
    if (e == cudaSuccess && !b->h_loff_full) {
        e = pinned_alloc(&b->pin_loff, (b->hv.max_reads + 1) * 4, "pinned line_off (on demand)");
        b->h_loff_full = (uint32_t*)b->pin_loff.p;
        if (e == cudaSuccess) { b->h_loff_full[0] = 0; b->h_line_off = b->h_loff_full; }
    }
    return e == cudaSuccess ? EXLR_OK : cuda_fail(e, "pinned result buffers");
}

extern "C" {

int exlr_abi_version(void) { return EXLR_ABI_VERSION; }

void exlr_params_default(exlr_params* p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->mapq = 1; p->exclude_flag = 1796; p->indel_min = 50; p->merge_min = 5; p->ins_clip_min = 1000;
    p->max_pct_overlap = 0.0; p->max_supp_alignm = 4;
}

int exlr_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDeviceCount"); return EXLR_ERR_CUDA; }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

const char* exlr_last_cuda_error(void) { return g_cuda_err; }

const char* exlr_strerror(int s)
{
    switch (s) {
    case EXLR_OK: return "ok";
    case EXLR_ERR_ARG: return "bad argument";
    case EXLR_ERR_CUDA: return "CUDA error or no usable sm_100 device (no CPU fallback exists)";
    case EXLR_ERR_NOMEM: return "out of memory";
    case EXLR_ERR_CAPACITY: return "batch exceeds its allocated capacity";
    case EXLR_ERR_STATE: return "call sequence error";
    case EXLR_ERR_TID: return "record passes the filters but has no valid reference id (reference panics in contig())";
    case EXLR_ERR_CIGAR_OP: return "unknown CIGAR operation code (reference panics)";
    case EXLR_ERR_SA_FIELDS: return "SA entry with fewer than 6 fields (reference panics)";
    case EXLR_ERR_SA_POS: return "SA position is not an integer (reference panics)";
    case EXLR_ERR_SA_STRAND: return "SA strand is not + or - (reference panics)";
    case EXLR_ERR_SA_CIGAR: return "SA CIGAR is malformed or outside [0-9MIDNSHP=X]";
    case EXLR_ERR_SA_MAPQ: return "SA mapq is not a u8 (reference panics)";
    case EXLR_ERR_SA_NM: return "SA NM is not an integer (reference panics)";
    case EXLR_ERR_MERGE_DOMAIN: return "more than two indel events with merge_min reaching across them: the merge loop of the reference indexes out of bounds and panics";
    case EXLR_ERR_SPLIT_COUNT: return "more than 2^24 segments in one record";
    case EXLR_ERR_BGZF: return "corrupt BGZF block (bad header, or a deflate stream that does not inflate to its ISIZE)";
    case EXLR_ERR_BAM_RECORD: return "corrupt BAM record (block_size or auxiliary data run past the record)";
    case EXLR_ERR_TEXT_CAPACITY: return "formatted lines exceed the batch's text buffer; use exlr_wait + exlr_format_lines";
    default: return "unknown status";
    }
}

int exlr_create(const exlr_params* p, int device, const char* const* ref_names, int n_ref, exlr_ctx** out)
{
    if (!p || !out || n_ref < 0 || (n_ref > 0 && !ref_names)) return EXLR_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    CK(traced("driver init + device count", 0, [&] { return cudaGetDeviceCount(&ndev); }));
    if (device < 0 || device >= ndev) return EXLR_ERR_ARG;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        snprintf(g_cuda_err, sizeof g_cuda_err, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return EXLR_ERR_CUDA;
    }
    CK(traced("context (cudaSetDevice + cudaFree(0))", 0, [&] { cudaError_t e = cudaSetDevice(device); return e == cudaSuccess ? cudaFree(nullptr) : e; }));
    int sms = 0;
    CK(traced("module load + kernel attributes", 0, [&] { return configure_kernels(device, &sms); }));
    exlr_ctx* c = new (std::nothrow) exlr_ctx();
    if (!c) return EXLR_ERR_NOMEM;
    c->device = device; c->params = *p; c->n_ref = n_ref; c->sms = sms;
    DevParams& d = c->dparams;
    d.mapq = p->mapq; d.exclude_flag = p->exclude_flag; d.exclude_secondary = p->exclude_secondary;
    d.exclude_unmapped = p->exclude_unmapped; d.split_only = p->split_only; d.indel_min = p->indel_min;
    d.merge_min = p->merge_min; d.ins_clip_min = p->ins_clip_min; d.max_pct_overlap = p->max_pct_overlap;
    d.max_supp_alignm = p->max_supp_alignm;
    c->far_mode = (uint64_t)p->merge_min > 2ull * p->indel_min;
    std::vector<uint32_t> off(n_ref + 1, 0);
    std::string bytes;
    for (int i = 0; i < n_ref; i++) {
        std::string s = ref_names[i] ? ref_names[i] : "";
        if (s.rfind("chr", 0) == 0) s = s.substr(3);
        c->ref_stripped.push_back(s);
        off[i] = (uint32_t)bytes.size();
        bytes += s;
    }
    off[n_ref] = (uint32_t)bytes.size();
    cudaError_t e = cudaMalloc(&c->d_ref_bytes, bytes.size() + 16);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_ref_off, off.size() * 4);
    if (e == cudaSuccess && !bytes.empty()) e = cudaMemcpy(c->d_ref_bytes, bytes.data(), bytes.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(c->d_ref_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_origin);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_origin, 0);
    if (e != cudaSuccess) { exlr_destroy(c); return cuda_fail(e, "reference name table"); }
    *out = c;
    return EXLR_OK;
}

void exlr_destroy(exlr_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_ref_bytes); cudaFree(c->d_ref_off);
    if (c->ev_origin) cudaEventDestroy(c->ev_origin);
    delete c;
}

int exlr_set_option(exlr_ctx* c, int option, int64_t value)
{
    if (!c) return EXLR_ERR_ARG;
    switch (option) {
    case EXLR_OPT_CIGAR_KERNEL: if (value < 0 || value > 3) return EXLR_ERR_ARG; c->cigar_kernel = (int)value; return EXLR_OK;
    case EXLR_OPT_READS_PER_CTA: if (value < 0 || value > 128) return EXLR_ERR_ARG; c->reads_per_cta = (uint32_t)value; return EXLR_OK;
    case EXLR_OPT_OVERLAP: if (value < 0 || value > 2) return EXLR_ERR_ARG; c->overlap = (int)value; return EXLR_OK;
    case EXLR_OPT_DEVICE_FORMAT: c->device_format = value != 0; return EXLR_OK;
    case EXLR_OPT_VERBOSE_TEXT: c->verbose_text = value != 0; return EXLR_OK;
    case EXLR_OPT_K3_FOLD: c->k3_fold = value != 0; return EXLR_OK;
    case EXLR_OPT_GRAPH: c->graph = value != 0; return EXLR_OK;
    case EXLR_OPT_WC_INPUT: c->wc_input = value != 0; return EXLR_OK;
    case EXLR_OPT_K0_WALK: c->k0_walk = value != 0; return EXLR_OK;
    case EXLR_OPT_BGZF_CRC: c->bgzf_crc = value != 0; return EXLR_OK;
    case EXLR_OPT_LONG_RECORDS: if (value < 0 || value > 3) return EXLR_ERR_ARG; c->long_records = (int)value; return EXLR_OK;
    case EXLR_OPT_K1A_CTAS_PER_SM: if (value < 1 || value > 8) return EXLR_ERR_ARG; c->k1a_ctas = (int)value; return EXLR_OK;
    case EXLR_OPT_TRACE: if (value < 0 || value > 7) return EXLR_ERR_ARG; c->trace = (int)value; return EXLR_OK;
    case EXLR_OPT_STAGE_TIMING: c->stage_timing = value != 0; return EXLR_OK;
    case EXLR_OPT_K1_WAVES: if (value < 1 || value > 16) return EXLR_ERR_ARG; c->k1_waves = (int)value; return EXLR_OK;
    case EXLR_OPT_K1_CTAS_PER_SM: if (value < 0 || value > 4) return EXLR_ERR_ARG; c->k1_ctas = (int)value; return EXLR_OK;
    default: return EXLR_ERR_ARG;
    }
}

void exlr_batch_free(exlr_batch* b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (auto& e : b->ev) if (e) cudaEventDestroy(e);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_k1_begin) cudaEventDestroy(b->ev_k1_begin);
    if (b->ev_k1_mid) cudaEventDestroy(b->ev_k1_mid);
    if (b->ev_k1_end) cudaEventDestroy(b->ev_k1_end);
    if (b->stream2) cudaStreamDestroy(b->stream2);
    if (b->stream) cudaStreamDestroy(b->stream);
    for (auto& e : b->ev_bam) if (e) cudaEventDestroy(e);
    if (b->ev_tail_copied) cudaEventDestroy(b->ev_tail_copied);
    cudaFree(b->d_bam); pinned_free(&b->pin_comp); cudaFreeHost(b->h_blocks); cudaFreeHost(b->h_btab); cudaFreeHost(b->h_bctrl);
    if (b->gexec) cudaGraphExecDestroy(b->gexec);
    cudaFree(b->d_slab); cudaFree(b->d_evslab);
    pinned_free(&b->pin_slab); cudaFreeHost(b->h_out); pinned_free(&b->pin_events); pinned_free(&b->pin_text); pinned_free(&b->pin_loff);
    delete b;
}

static int batch_alloc_impl(exlr_ctx* c, uint64_t max_reads, uint64_t max_ops, uint64_t max_sa_bytes, uint64_t max_events, bool host_inputs, bool verbose_text, exlr_batch** out)   // host_inputs = false: a BAM batch (no pinned input views, lean pinned results)
{
    if (!c || !out) return EXLR_ERR_ARG;
    *out = nullptr;
    if (max_reads == 0 || max_reads >= 0xfffffff0ull || max_sa_bytes >= 0x7ffffff0ull || max_events >= 0xfff00000ull) return EXLR_ERR_ARG;
    if (max_events == 0) max_events = 2 * max_reads + 1024;
    CK(cudaSetDevice(c->device));
    exlr_batch* b = new (std::nothrow) exlr_batch();
    if (!b) return EXLR_ERR_NOMEM;
    b->ctx = c;
    const size_t R = max_reads, A = 256;
    // ---- pinned host input slab
    size_t ho = 0;
    auto hcarve = [&](size_t bytes) { size_t at = ho; ho = align_up(ho + bytes, A); return at; };
    const size_t h_cigar = hcarve((max_ops + 4) * 4), h_coff = hcarve((R + 1) * 8), h_pos = hcarve(R * 4), h_tid = hcarve(R * 4),
                 h_flag = hcarve(R * 2), h_mapq = hcarve(R), h_kind = hcarve(R), h_soff = hcarve((R + 1) * 4), h_sab = hcarve(max_sa_bytes + 16);
    cudaError_t e = cudaSuccess;
    b->hv.max_reads = max_reads; b->hv.max_ops = max_ops; b->hv.max_sa_bytes = max_sa_bytes; b->hv.max_events = max_events;
    if (host_inputs) {                         // (a BAM batch gets its records from the device-side decoder: no pinned input views)
        // (write-combined: the host only ever writes these views front to back; EXLR_OPT_WC_INPUT, off by default -- the host formatter
        // and check_sizes read a few words of them back, which is slow on write-combined memory but rare)
        if (c->wc_input) { e = host_alloc(&b->pin_slab.p, ho, cudaHostAllocWriteCombined, "pinned input views (write-combined)"); }
        else e = pinned_alloc(&b->pin_slab, ho, "pinned input views");
        b->h_slab = b->pin_slab.p;
        if (e != cudaSuccess) { exlr_batch_free(b); return cuda_fail(e, "cudaHostAlloc(inputs)"); }
        char* hs = (char*)b->h_slab;
        b->hv.cigar = (uint32_t*)(hs + h_cigar); b->hv.cigar_off = (uint64_t*)(hs + h_coff); b->hv.pos = (int32_t*)(hs + h_pos);
        b->hv.tid = (int32_t*)(hs + h_tid); b->hv.flag = (uint16_t*)(hs + h_flag); b->hv.mapq = (uint8_t*)(hs + h_mapq);
        b->hv.sa_kind = (uint8_t*)(hs + h_kind); b->hv.sa_off = (uint32_t*)(hs + h_soff); b->hv.sa_bytes = (uint8_t*)(hs + h_sab);
        b->hv.cigar_off[0] = 0; b->hv.sa_off[0] = 0;
    }
    // ---- pinned host output slab
    size_t oo = 0;
    auto ocarve = [&](size_t bytes) { size_t at = oo; oo = align_up(oo + bytes, A); return at; };
    b->lean_host = !host_inputs;
    const size_t o_ctrl = ocarve(sizeof(Ctrl)), o_loff = ocarve(b->lean_host ? 4 : (R + 1) * 4);   // (lean: line_off[0] only, see ensure_fetch_buffers)
    e = host_alloc(&b->h_out, oo, cudaHostAllocMapped, "pinned header+line_off");
    if (e != cudaSuccess) { exlr_batch_free(b); return cuda_fail(e, "cudaHostAlloc(outputs)"); }
    b->h_ctrl = (Ctrl*)((char*)b->h_out + o_ctrl); b->h_line_off = (uint32_t*)((char*)b->h_out + o_loff);
    e = cudaHostGetDevicePointer((void**)&b->h_ctrl_dev, b->h_ctrl, 0);
    if (e != cudaSuccess) { exlr_batch_free(b); return cuda_fail(e, "cudaHostGetDevicePointer"); }
    b->dv.host_ctrl = b->h_ctrl_dev;
    b->device_format = c->device_format != 0; b->verbose_text = verbose_text;
    // ---- device slab
    const uint32_t tiles = scan_tiles((uint32_t)R);
    const bool need_pool = c->params.max_supp_alignm + 1 > (uint64_t)kLocalSegs;
    const size_t pool_cap = need_pool ? (max_sa_bytes / 10 + R + 64) : 0;
    size_t dof = 0;
    auto dcarve = [&](size_t bytes) { size_t at = dof; dof = align_up(dof + bytes, A); return at; };
    const size_t bits_bytes = align_up((R + 31) / 32 * 4, 16);
    const size_t d_ctrl = dcarve(sizeof(Ctrl) + (size_t)tiles * 16 + bits_bytes);   // ctrl | scan_a | scan_b | dirty_bits : one memset (scan_c lives with the event buffers)
    const size_t d_cigar = dcarve((max_ops + 4) * 4 + 16), d_coff = dcarve((R + 1) * 8), d_pos = dcarve(R * 4), d_tid = dcarve(R * 4),
                 d_flag = dcarve(R * 2), d_mapq = dcarve(R), d_kind = dcarve(R), d_soff = dcarve((R + 1) * 4), d_sab = dcarve(max_sa_bytes + 16),
                 d_k1 = dcarve(R * 8), d_tcnt = dcarve(R * 4), d_slist = dcarve((max_ops / 512 + 16) * 4), d_ssum = dcarve((max_ops / 512 + 16) * 4), d_sflag = dcarve(max_ops / 512 + 16), d_llist = dcarve(R * 4), d_far = dcarve(R * 8), d_shlist = dcarve(R * 4), d_wlist = dcarve(R * 4), d_csa = dcarve(R * 4), d_list = dcarve(R * 4), d_base = dcarve(R * 4), d_sum = dcarve(R * sizeof(SaSum)),
                 d_pool = dcarve(pool_cap * sizeof(Seg)), d_dbg = dcarve(8192 * 32), d_loff = dcarve((R + 1) * 4);
    e = dev_alloc(&b->d_slab, dof, "device batch");
    if (e != cudaSuccess) { exlr_batch_free(b); return cuda_fail(e, "cudaMalloc(batch)"); }
    char* ds = (char*)b->d_slab;
    DevBatch& v = b->dv;
    v.ctrl = (Ctrl*)(ds + d_ctrl);
    v.scan_a = (unsigned long long*)(ds + d_ctrl + sizeof(Ctrl)); v.scan_b = v.scan_a + tiles;
    v.dirty_bits = (uint32_t*)(v.scan_b + tiles); v.step_list = (uint32_t*)(ds + d_slist); v.step_sum = (uint32_t*)(ds + d_ssum); v.step_flag = (uint8_t*)(ds + d_sflag); v.long_list = (uint32_t*)(ds + d_llist); v.far_list = (uint2*)(ds + d_far); v.short_list = (uint32_t*)(ds + d_shlist); v.warp_list = (uint32_t*)(ds + d_wlist);
    b->ctrl_bytes = sizeof(Ctrl) + (size_t)tiles * 16 + bits_bytes;
    v.cigar = (uint32_t*)(ds + d_cigar); v.cigar_off = (unsigned long long*)(ds + d_coff); v.pos = (int32_t*)(ds + d_pos);
    v.tid = (int32_t*)(ds + d_tid); v.flag = (uint16_t*)(ds + d_flag); v.mapq = (uint8_t*)(ds + d_mapq); v.sa_kind = (uint8_t*)(ds + d_kind);
    v.sa_off = (uint32_t*)(ds + d_soff); v.sa_bytes = (uint8_t*)(ds + d_sab);
    v.ref_bytes = c->d_ref_bytes; v.ref_off = c->d_ref_off; v.n_ref = c->n_ref;
    v.tile_cnt = (uint32_t*)(ds + d_tcnt);
    v.k1 = (uint2*)(ds + d_k1); v.csa = (uint32_t*)(ds + d_csa); v.sa_list = (uint32_t*)(ds + d_list); v.sa_base = (uint32_t*)(ds + d_base);
    v.sa_sum = (SaSum*)(ds + d_sum);
    v.seg_pool = (Seg*)(ds + d_pool); v.seg_pool_cap = (uint32_t)pool_cap;
    b->d_dbg = (unsigned long long*)(ds + d_dbg); v.dbg = nullptr;
    v.line_off = (uint32_t*)(ds + d_loff);
    v.n_reads = 0;
    { const int rc = alloc_event_buffers(b, max_events); if (rc) { exlr_batch_free(b); return rc; } }
    // the SA branch (kernels 0, 3a, 3b) is the longer chain: its stream gets the higher priority so its CTAs are placed first
    // whenever kernel 1 (stream2) frees a slot
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    e = cudaStreamCreateWithPriority(&b->stream, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&b->stream2, cudaStreamNonBlocking, prio_lo);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev_fork);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev_k1_begin);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev_k1_mid);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev_k1_end);
    for (int i = 0; i < EV_COUNT && e == cudaSuccess; i++) e = cudaEventCreate(&b->ev[i]);
    if (e != cudaSuccess) { exlr_batch_free(b); return cuda_fail(e, "stream/event creation"); }
    *out = b;
    return EXLR_OK;
}

int exlr_batch_alloc(exlr_ctx* c, uint64_t max_reads, uint64_t max_ops, uint64_t max_sa_bytes, uint64_t max_events, exlr_batch** out)
{
    return batch_alloc_impl(c, max_reads, max_ops, max_sa_bytes, max_events, true, false, out);
}

int exlr_batch_grow(exlr_batch* b, uint64_t max_events)
{
    if (!b) return EXLR_ERR_ARG;
    if (max_events <= b->hv.max_events) return EXLR_OK;
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaStreamSynchronize(b->stream));
    if (b->stream2) CK(cudaStreamSynchronize(b->stream2));
    free_event_buffers(b);
    drop_graph(b);
    b->submitted = false; b->have_timing = false;
    return alloc_event_buffers(b, max_events);
}

int exlr_batch_get_views(exlr_batch* b, exlr_batch_views* v)
{
    if (!b || !v) return EXLR_ERR_ARG;
    if (b->is_bam) return EXLR_ERR_STATE;      // a BAM batch has no host-side record views
    *v = b->hv;
    return EXLR_OK;
}

static int check_sizes(exlr_batch* b, uint64_t n_reads)
{
    if (n_reads > b->hv.max_reads) return EXLR_ERR_CAPACITY;
    if (b->hv.cigar_off[0] != 0 || b->hv.sa_off[0] != 0) return EXLR_ERR_ARG;
    if (b->hv.cigar_off[n_reads] > b->hv.max_ops || b->hv.sa_off[n_reads] > b->hv.max_sa_bytes) return EXLR_ERR_CAPACITY;
    return EXLR_OK;
}

static int copy_inputs(exlr_batch* b, uint64_t n)
{
    const exlr_batch_views& h = b->hv; DevBatch& d = b->dv; cudaStream_t st = b->stream;
    const uint64_t ops = h.cigar_off[n], sab = h.sa_off[n];
    b->h2d_bytes = ops * 4 + sab + n * 24 + 12;
    if (ops) CK(cudaMemcpyAsync((void*)d.cigar, h.cigar, ops * 4, cudaMemcpyHostToDevice, st));
    if (n == h.max_reads) {
        // a full batch: cigar_off .. sa_off are carved back to back with the same relative offsets on both sides, one copy moves them all
        const size_t bytes = (size_t)((const char*)(h.sa_off + n + 1) - (const char*)h.cigar_off);
        CK(cudaMemcpyAsync((void*)d.cigar_off, h.cigar_off, bytes, cudaMemcpyHostToDevice, st));
        if (sab) CK(cudaMemcpyAsync((void*)d.sa_bytes, h.sa_bytes, sab, cudaMemcpyHostToDevice, st));
        return EXLR_OK;
    }
    CK(cudaMemcpyAsync((void*)d.cigar_off, h.cigar_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)d.pos, h.pos, n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)d.tid, h.tid, n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)d.flag, h.flag, n * 2, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)d.mapq, h.mapq, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)d.sa_kind, h.sa_kind, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((void*)d.sa_off, h.sa_off, (n + 1) * 4, cudaMemcpyHostToDevice, st));
    if (sab) CK(cudaMemcpyAsync((void*)d.sa_bytes, h.sa_bytes, sab, cudaMemcpyHostToDevice, st));
    return EXLR_OK;
}

// batches whose mean CIGAR is longer than this run kernel 1c behind kernel 1b
static constexpr uint64_t kLongRecordMeanOps = 512;

static uint32_t auto_rpc(uint64_t n_reads, uint64_t n_ops)
{
    // kernel 1 scans 2048 ops per step: aim just under two steps of CIGAR per CTA for short-read batches (full
    // 128-record CTAs), one record per CTA once records are longer than that
    const uint64_t mean = n_reads ? (n_ops + n_reads - 1) / n_reads : 1;
    uint64_t rpc = 3900 / (mean ? mean : 1);
    if (rpc < 1) rpc = 1;
    if (rpc > 128) rpc = 128;
    return (uint32_t)rpc;
}

// How many events / text bytes a submit copies back behind its kernels without knowing the real count: a little more than the
// last waited batch of this context produced.  exlr_wait then needs one synchronisation; what the guess missed is fetched after it.
static uint64_t guess_with_margin(uint64_t hint, uint64_t first_guess, uint64_t cap)
{
    const uint64_t g = hint ? hint + hint / 8 + 1024 : first_guess;
    return g < cap ? g : cap;
}

static void plan_step(exlr_batch* b, StepPlan* p)
{
    exlr_ctx* c = b->ctx;
    memset(p, 0, sizeof(*p));
    p->n_reads = b->n_reads; p->n_ops = b->n_ops; p->events = b->dv.events; p->trace = c->trace;
    p->overlap = c->overlap && !c->params.split_only;
    p->fork_late = p->overlap && c->overlap == 2;
    p->variant = c->cigar_kernel == 1 ? 1 : 0;
    if (!c->params.split_only) {
        p->rpc = c->reads_per_cta ? c->reads_per_cta : auto_rpc(b->n_reads, b->n_ops);
        // Events (I/D >= indel_min) are sparse in HiFi and ONT batches alike (~10 % of the records hold one), so by default the
        // CIGAR stream is screened at streaming speed (kernel 1a) and only the records around an event candidate are scanned
        // (kernels 1b, 1c).  A batch where most 512-op steps held a candidate (e.g. -i 1) is better off with the flat scan of
        // everything: after such a batch the next few run unscreened, then the screen is tried again.
        // (kernel 1a indexes the CIGAR array by 32-bit vector numbers)
        bool want = c->cigar_kernel == 3;
        if (c->cigar_kernel == 0) { if (c->skip_screen.load() > 0) c->skip_screen.fetch_sub(1); else want = true; }
        p->screened = want && b->n_ops < (1ull << 33) && b->n_ops > 0;
        // Batches of long records (ONT-like): kernel 1a also leaves per-step sums, and the long records behind it are resolved by
        // kernel 1d from those sums (EXLR_OPT_LONG_RECORDS=2: by kernel 1c, the flat block scan of the listed records).  In a
        // batch of short records the odd long one is scanned by a warp of kernel 1b, and the chain is one launch shorter.
        p->long_batch = p->screened && (c->long_records ? c->long_records >= 2 : b->n_ops / b->n_reads > kLongRecordMeanOps);
        p->two_level = p->long_batch && c->long_records != 2;
    }
    const uint32_t mean_ops = (uint32_t)(b->n_reads ? b->n_ops / b->n_reads : 0);
    p->fold = k3_fold(mean_ops) && c->k3_fold;                        // short CIGARs: kernel 3b walks the SA records' own CIGARs itself (off by default)
    p->k0_walk = !p->fold && k3_fold(mean_ops) && c->k0_walk;         // short CIGARs: kernel 0 does, for the records it lists
    p->far = c->far_mode;
    p->formatted = b->dv.text_off != nullptr;
}

// The step itself: two memsets and the kernels, on the batch's two streams.  `timed`: CUDA events between the kernels
// (exlr_timing per stage); never while the step is being captured into a graph.
static int enqueue_step(exlr_batch* b, const StepPlan& p, bool timed)
{
    exlr_ctx* c = b->ctx; cudaStream_t st = b->stream; DevBatch& d = b->dv;
    b->launches = 0;
    d.dbg = c->trace ? b->d_dbg : nullptr; d.dbg_sel = (uint32_t)c->trace;
    if (c->trace) CK(cudaMemsetAsync(b->d_dbg, 0, 8192 * 32, st));
    CK(cudaMemsetAsync(d.ctrl, 0, b->ctrl_bytes, st));
    if (b->scan_c_bytes) CK(cudaMemsetAsync(d.scan_c, 0, b->scan_c_bytes, st));
    // kernel 1 needs nothing from kernel 0, so it runs on a second stream beside the SA branch (0 -> 3a -> 3b); 4a joins them
    d.prim_slots = 0; d.capt_log2 = 0;
    d.hc = HostCfg{c->sms, c->k1_ctas ? c->k1_ctas : (p.overlap ? 3 : 4), c->k1a_ctas, c->k1_waves};
    d.k1_gated = p.screened ? 1u : 0u;
    d.qnames = b->is_bam ? b->db.qnames : nullptr; d.qname_off = b->is_bam ? b->db.qname_off : nullptr;
    d.verbose = b->is_bam && b->verbose_text ? 1u : 0u;
    if (!c->params.split_only) {
        uint32_t n_tiles = 0;
        plan_k1(d, p.screened ? 1 : p.variant, p.rpc, &n_tiles);         // screened: raw events all go to the atomically allocated region
        if (p.overlap && !p.fork_late) CK(cudaEventRecord(b->ev_fork, st));   // after the memset
    }
    launch_k0(d, c->dparams, p.k0_walk, st); b->launches++;
    if (p.fork_late) CK(cudaEventRecord(b->ev_fork, st));                 // EXLR_OPT_OVERLAP = 2: the CIGAR path starts once kernel 0 is done
    if (timed) CK(cudaEventRecord(b->ev[EV_K0], st));
    if (!c->params.split_only) {
        cudaStream_t s1 = p.overlap ? b->stream2 : st;
        if (p.overlap) CK(cudaStreamWaitEvent(s1, b->ev_fork, 0));
        if (timed) CK(cudaEventRecord(b->ev_k1_begin, s1));
        if (p.screened) {
            launch_k1a(d, c->dparams, b->n_ops, p.two_level, s1); b->launches++;
            if (timed) CK(cudaEventRecord(b->ev_k1_mid, s1));
            launch_k1b(d, c->dparams, b->n_ops, p.long_batch, s1); b->launches += 2;
            if (p.long_batch) { if (p.two_level) launch_k1d(d, c->dparams, s1); else launch_k1c(d, c->dparams, s1); b->launches++; }
        } else {
            launch_k1(d, c->dparams, p.variant, p.rpc, s1); b->launches++;
        }
        CK(cudaEventRecord(b->ev_k1_end, s1));
    }
    if (timed) CK(cudaEventRecord(b->ev[EV_K1], st));
    if (!p.fold && !p.k0_walk) { launch_k3a(d, c->dparams, (uint32_t)(b->n_reads ? b->n_ops / b->n_reads : 0), st); b->launches++; }
    if (timed) CK(cudaEventRecord(b->ev[EV_K3A], st));
    launch_k3b(d, c->dparams, p.fold, st); b->launches++;
    if (timed) CK(cudaEventRecord(b->ev[EV_K3B], st));
    if (p.overlap) CK(cudaStreamWaitEvent(st, b->ev_k1_end, 0));
    launch_k4a(d, c->dparams, p.far, st); b->launches++;
    if (timed) CK(cudaEventRecord(b->ev[EV_K4A], st));
    launch_k4b(d, c->dparams, p.far, st); b->launches++;
    if (p.formatted) { launch_k5(d, st); b->launches += 2; }       // (no event in between: 5a is placed while 4b drains)
    CK(cudaGetLastError());
    return EXLR_OK;
}

static int run_kernels(exlr_batch* b, bool prefetch_results)
{
    exlr_ctx* c = b->ctx; cudaStream_t st = b->stream; DevBatch& d = b->dv;
    StepPlan plan;
    plan_step(b, &plan);
    b->screened = plan.screened; b->formatted = plan.formatted;
    b->stage_timed = c->stage_timing != 0;
    // A caller that submits the same shape again and again (a resident batch re-run, a ring of equal sub-batches) gets the step
    // as ONE CUDA graph launch from the third time on: the chain of 5-9 small kernels is then scheduled by the device, not
    // paced by this thread's launch calls (which jitter when several processes share the host).  The graph keeps the two-stream
    // fork/join and the programmatic-dependent-launch edges.  Varying shapes (a streamed BAM) launch directly, as before.
    const bool graph_ok = c->graph && !c->stage_timing;
    if (graph_ok && b->gexec && plan == b->gplan) {
        CK(cudaGraphLaunch(b->gexec, st));
        b->launches = b->glaunches;
    } else {
        bool captured = false;
        if (graph_ok && b->have_last_plan && plan == b->last_plan) {
            drop_graph(b);
            cudaGraph_t g = nullptr;
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int rc = enqueue_step(b, plan, false);
                const cudaError_t e = cudaStreamEndCapture(st, &g);
                if (rc == EXLR_OK && e == cudaSuccess && g && cudaGraphInstantiate(&b->gexec, g, 0) == cudaSuccess) {
                    b->gplan = plan; b->glaunches = b->launches; captured = true;
                } else { b->gexec = nullptr; c->graph = 0; cudaGetLastError(); }     // this driver cannot capture the step: never try again
                if (g) cudaGraphDestroy(g);
            } else { c->graph = 0; cudaGetLastError(); }
            if (captured) CK(cudaGraphLaunch(b->gexec, st));
        }
        if (!captured) { const int rc = enqueue_step(b, plan, c->stage_timing != 0); if (rc) return rc; }
    }
    b->last_plan = plan; b->have_last_plan = true;
    b->far_ran = plan.far;
    CK(cudaEventRecord(b->ev[EV_K4B], st));
    b->d2h_events = 0; b->d2h_text = 0; b->have_line_off = false; b->d2h_bytes = sizeof(Ctrl);
    if (prefetch_results) {
        // results follow the kernels on the same stream, sized by a guess (the counts live on the device): lines when the batch
        // formats them (exlr_wait_text), else line offsets + events (exlr_wait)
        if (b->formatted) {
            b->d2h_text = guess_with_margin(c->text_hint.load(), b->n_reads * 16 + 65536, std::min<uint64_t>(d.text_cap, b->h_text_cap));
            if (b->d2h_text) CK(cudaMemcpyAsync(b->h_text, d.text, b->d2h_text, cudaMemcpyDeviceToHost, st));
            b->d2h_bytes += b->d2h_text;
        } else {
            { const int rc = ensure_fetch_buffers(b); if (rc) return rc; }
            CK(cudaMemcpyAsync(b->h_line_off, d.line_off, (b->n_reads + 1) * 4, cudaMemcpyDeviceToHost, st));
            b->have_line_off = true;
            b->d2h_events = guess_with_margin(c->ev_hint.load(), b->n_reads / 4 + 4096, d.max_events);
            CK(cudaMemcpyAsync(b->h_events, d.events, b->d2h_events * sizeof(exlr_event), cudaMemcpyDeviceToHost, st));
            b->d2h_bytes += (b->n_reads + 1) * 4 + b->d2h_events * sizeof(exlr_event);
        }
    }
    CK(cudaEventRecord(b->ev[EV_D2H], st));
    CK(cudaGetLastError());
    return EXLR_OK;
}

int exlr_submit(exlr_batch* b, uint64_t n_reads)
{
    if (!b) return EXLR_ERR_ARG;
    if (b->is_bam) return EXLR_ERR_STATE;
    int rc = check_sizes(b, n_reads);
    if (rc) return rc;
    CK(cudaSetDevice(b->ctx->device));
    b->n_reads = n_reads; b->n_ops = b->hv.cigar_off[n_reads]; b->dv.n_reads = (uint32_t)n_reads;
    b->resident_uploaded = false; b->have_timing = false;
    if (n_reads == 0) { memset(b->h_ctrl, 0, sizeof(Ctrl)); b->h_line_off[0] = 0; b->submitted = true; return EXLR_OK; }
    CK(cudaEventRecord(b->ev[EV_START], b->stream));
    rc = copy_inputs(b, n_reads);
    if (rc) return rc;
    CK(cudaEventRecord(b->ev[EV_H2D], b->stream));
    rc = run_kernels(b, true);
    if (rc) return rc;
    b->submitted = true; b->have_timing = true;
    return EXLR_OK;
}

int exlr_upload(exlr_batch* b, uint64_t n_reads)
{
    if (!b) return EXLR_ERR_ARG;
    if (b->is_bam) return EXLR_ERR_STATE;
    int rc = check_sizes(b, n_reads);
    if (rc) return rc;
    if (n_reads == 0) return EXLR_ERR_ARG;
    CK(cudaSetDevice(b->ctx->device));
    b->n_reads = n_reads; b->n_ops = b->hv.cigar_off[n_reads]; b->dv.n_reads = (uint32_t)n_reads;
    rc = copy_inputs(b, n_reads);
    if (rc) return rc;
    CK(cudaStreamSynchronize(b->stream));
    b->resident_uploaded = true; b->submitted = false;
    return EXLR_OK;
}

int exlr_submit_resident(exlr_batch* b)
{
    if (!b) return EXLR_ERR_ARG;
    if (!b->resident_uploaded) return EXLR_ERR_STATE;
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaEventRecord(b->ev[EV_START], b->stream));
    CK(cudaEventRecord(b->ev[EV_H2D], b->stream));
    int rc = run_kernels(b, false);
    if (rc) return rc;
    b->submitted = true; b->have_timing = true;
    return EXLR_OK;
}

static int status_of_rank(uint32_t rank)
{
    switch (rank) {
    case RANK_TID: return EXLR_ERR_TID; case RANK_CIGAR_OP: return EXLR_ERR_CIGAR_OP; case RANK_SA_FIELDS: return EXLR_ERR_SA_FIELDS;
    case RANK_SA_POS: return EXLR_ERR_SA_POS; case RANK_SA_STRAND: return EXLR_ERR_SA_STRAND; case RANK_SA_CIGAR: return EXLR_ERR_SA_CIGAR;
    case RANK_SA_MAPQ: return EXLR_ERR_SA_MAPQ; case RANK_SA_NM: return EXLR_ERR_SA_NM; case RANK_MERGE_DOMAIN: return EXLR_ERR_MERGE_DOMAIN;
    default: return EXLR_ERR_SPLIT_COUNT;
    }
}

static int finish(exlr_batch* b, exlr_result* res, bool fetch)
{
    if (!b || !res) return EXLR_ERR_ARG;
    if (!b->submitted) return EXLR_ERR_STATE;
    memset(res, 0, sizeof(*res));
    res->n_reads = b->n_reads; res->n_ops = b->n_ops; res->err_read = 0xffffffffu;
    res->events = b->h_events; res->line_off = b->h_line_off;
    if (b->n_reads == 0) { res->status = EXLR_OK; return EXLR_OK; }
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaStreamSynchronize(b->stream));
    if (b->h_ctrl->need_far && !b->far_ran) {
        // a record for the literal >2 merge loop in a batch whose parameters rule it out short of a u32 wrap (see k4a_line_scan):
        // the tail of the pipeline runs again with the loop compiled in
        exlr_ctx* cx = b->ctx; cudaStream_t st = b->stream;
        launch_reset_tail(b->dv, st);
        launch_k4a(b->dv, cx->dparams, true, st);
        launch_k4b(b->dv, cx->dparams, true, st);
        if (b->formatted) launch_k5(b->dv, st);
        b->far_ran = true; b->launches += b->formatted ? 5 : 3;
        b->d2h_events = 0; b->d2h_text = 0; b->have_line_off = false;      // what was copied back before is stale
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
    }
    const Ctrl& c = *b->h_ctrl;
    if (b->screened && (uint64_t)c.n_flagged * 2 > (uint64_t)k1a_steps(b->n_ops)) b->ctx->skip_screen.store(8);   // event-dense: see run_kernels
    res->n_events = c.n_events; res->n_kept = c.n_kept; res->n_sa_reads = c.n_sa; res->n_cap_dropped = c.n_dropped;
    if (c.overflow) {
        // n_events = the max_events that would have sufficed (all three counters keep counting past the capacity)
        uint64_t need = c.n_events; if (c.n_saev > need) need = c.n_saev;
        if (c.n_raw) { const uint64_t r = (uint64_t)b->dv.raw_cap + 2ull * c.n_raw; if (r > need) need = r; }
        res->status = EXLR_ERR_CAPACITY; res->n_events = need; return res->status;
    }
    b->ctx->ev_hint.store(c.n_events); b->ctx->text_hint.store(c.text_bytes);
    if (fetch) {
        { const int rc = ensure_fetch_buffers(b); if (rc) return rc; res->events = b->h_events; res->line_off = b->h_line_off; }
        // usually everything is here already (copied behind the kernels by exlr_submit); fetch what the guess missed
        bool more = false;
        if (!b->have_line_off) {
            CK(cudaMemcpyAsync(b->h_line_off, b->dv.line_off, (b->n_reads + 1) * 4, cudaMemcpyDeviceToHost, b->stream));
            b->have_line_off = true; more = true; b->d2h_bytes += (b->n_reads + 1) * 4;
        }
        if (c.n_events > b->d2h_events) {
            const size_t at = (size_t)b->d2h_events, cnt = (size_t)c.n_events - at;
            CK(cudaMemcpyAsync(b->h_events + at, b->dv.events + at, cnt * sizeof(exlr_event), cudaMemcpyDeviceToHost, b->stream));
            b->d2h_events = c.n_events; more = true; b->d2h_bytes += cnt * sizeof(exlr_event);
        }
        if (more) CK(cudaStreamSynchronize(b->stream));
    }
    if (c.err_key) {
        const unsigned long long key = ~c.err_key;
        res->err_read = (uint32_t)(key >> 8);
        res->status = status_of_rank((uint32_t)(key & 0xff));
        res->n_err_lines = c.err_lines;
    }
    return res->status;
}

int exlr_wait(exlr_batch* b, exlr_result* res) { return finish(b, res, true); }

int exlr_wait_text(exlr_batch* b, exlr_result* res, const char** text, uint64_t* n_bytes)
{
    if (!text || !n_bytes) return EXLR_ERR_ARG;
    *text = nullptr; *n_bytes = 0;
    int rc = finish(b, res, false);
    if (rc != EXLR_OK && rc > -10) return rc;                           // no usable result at all
    if (!b->formatted || b->n_reads == 0) return b->n_reads == 0 ? rc : EXLR_ERR_STATE;
    const Ctrl& c = *b->h_ctrl;
    if (c.text_bytes > b->dv.text_cap) return EXLR_ERR_TEXT_CAPACITY;    // exlr_wait + exlr_format_lines still work
    uint64_t nb = c.text_bytes;
    if (rc <= -10) {
        // a record on which the reference panics: the lines of the records before it stand (its BufWriter is flushed on unwind),
        // and so do the lines the record itself had written by then
        uint32_t first_line = 0, off = 0;
        CK(cudaMemcpyAsync(&first_line, b->dv.line_off + res->err_read, 4, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
        CK(cudaMemcpyAsync(&off, b->dv.text_off + first_line + (uint32_t)res->n_err_lines, 4, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
        nb = off;
    }
    if (nb > b->h_text_cap) {                                            // lean batch: more text than its pinned buffer holds yet
        pinned_free(&b->pin_text); b->h_text = nullptr; b->d2h_text = 0;
        b->h_text_cap = std::min<uint64_t>(b->dv.text_cap, nb + nb / 2);
        const cudaError_t e = pinned_alloc(&b->pin_text, b->h_text_cap + 16, "pinned text (grown)");
        b->h_text = (char*)b->pin_text.p;
        if (e != cudaSuccess) { b->h_text_cap = 0; return cuda_fail(e, "pinned text"); }
    }
    if (nb > b->d2h_text) {                                              // what the copy behind the kernels did not cover
        const uint64_t at = b->d2h_text;
        CK(cudaMemcpyAsync(b->h_text + at, b->dv.text + at, nb - at, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
        b->d2h_text = nb; b->d2h_bytes += nb - at;
    }
    *text = b->h_text; *n_bytes = nb;
    return rc;
}
int exlr_wait_resident(exlr_batch* b, exlr_result* res) { return finish(b, res, false); }

int exlr_get_timing(exlr_batch* b, exlr_timing* t)
{
    if (!b || !t) return EXLR_ERR_ARG;
    if (!b->have_timing) return EXLR_ERR_STATE;
    memset(t, 0, sizeof(*t));
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaEventSynchronize(b->ev[EV_D2H]));
    CK(cudaEventElapsedTime(&t->h2d_ms, b->ev[EV_START], b->ev[EV_H2D]));
    if (b->stage_timed) {
        CK(cudaEventElapsedTime(&t->classify_ms, b->ev[EV_H2D], b->ev[EV_K0]));
        if (b->ctx->params.split_only) t->cigar_ms = 0.f; else CK(cudaEventElapsedTime(&t->cigar_ms, b->ev_k1_begin, b->ev_k1_end));
        if (b->screened && !b->ctx->params.split_only) CK(cudaEventElapsedTime(&t->screen_ms, b->ev_k1_begin, b->ev_k1_mid));
        CK(cudaEventElapsedTime(&t->sa_cigar_ms, b->ev[EV_K1], b->ev[EV_K3A]));
        CK(cudaEventElapsedTime(&t->sa_parse_ms, b->ev[EV_K3A], b->ev[EV_K3B]));
        CK(cudaEventElapsedTime(&t->scan_ms, b->ev[EV_K3B], b->ev[EV_K4A]));
        CK(cudaEventElapsedTime(&t->place_ms, b->ev[EV_K4A], b->ev[EV_K4B]));
    }
    CK(cudaEventElapsedTime(&t->kernels_ms, b->ev[EV_H2D], b->ev[EV_K4B]));
    CK(cudaEventElapsedTime(&t->d2h_ms, b->ev[EV_K4B], b->ev[EV_D2H]));
    t->launches = b->launches;
    t->h2d_bytes = b->h2d_bytes; t->d2h_bytes = b->d2h_bytes;
    return EXLR_OK;
}

// ---- BAM input decoded on the device ---------------------------------------------------------------------------------------
int exlr_bam_batch_alloc(exlr_ctx* c, uint64_t max_comp_bytes, uint32_t max_blocks, uint64_t max_tail_bytes, uint64_t max_events, exlr_batch** out)
{
    if (!c || !out || max_comp_bytes == 0 || max_blocks == 0 || max_blocks > 32000u || max_comp_bytes >= 0xf0000000ull || max_tail_bytes >= 0x40000000ull) return EXLR_ERR_ARG;
    // what a chunk of that many BGZF blocks (64 KB of BAM each at most) plus the previous chunk's tail can hold: every bound is
    // exact, so no chunk overflows its batch
    const uint64_t front_u = align_up(max_tail_bytes, 256);
    const uint64_t u_cap = (uint64_t)max_blocks * 65536ull + front_u;
    const uint64_t R = u_cap / 36 + 1, OPS = u_cap / 4 + 4, SAB = u_cap;
    if (max_events == 0) max_events = R / 8 + 65536;
    const bool fmt = c->device_format != 0;
    c->device_format = 1;                      // the lines of a BAM batch are always formatted on the device (exlr_wait_text)
    exlr_batch* b = nullptr;
    const int rc = batch_alloc_impl(c, R, OPS, SAB, max_events, false, c->verbose_text != 0, &b);
    c->device_format = fmt;
    if (rc) return rc;
    b->is_bam = true; b->max_comp = max_comp_bytes; b->max_blocks = max_blocks; b->u_cap = u_cap; b->front_u = front_u;
    cudaError_t e = pinned_alloc(&b->pin_comp, max_comp_bytes + 512, "pinned BGZF chunk");
    b->h_comp = (uint8_t*)b->pin_comp.p;
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&b->h_blocks, (size_t)max_blocks * sizeof(exlr_bgzf_block), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&b->h_btab, ((size_t)max_blocks + 1) * sizeof(BgzfBlock), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&b->h_bctrl, sizeof(BamCtrl), cudaHostAllocMapped);
    if (e != cudaSuccess) { exlr_batch_free(b); return cuda_fail(e, "cudaHostAlloc(BAM chunk)"); }
    memset(b->h_bctrl, 0, sizeof(BamCtrl));
    const uint32_t tiles = bam_scan_tiles((uint32_t)R);
    const size_t tab_n = (size_t)max_blocks + 1;               // entry 0: the previous chunk's tail, as one pseudo block
    size_t dof = 0;
    auto dcarve = [&](size_t bytes) { size_t at = dof; dof = align_up(dof + bytes, 256); return at; };
    const size_t d_ctrl = dcarve(sizeof(BamCtrl) + (size_t)tiles * 24), d_comp = dcarve(max_comp_bytes + 1024), d_tab = dcarve(tab_n * sizeof(BgzfBlock)),
                 d_u = dcarve(u_cap + 256), d_blk = dcarve(tab_n * 4 * 6), d_rec = dcarve(R * 4), d_per = dcarve(R * 4 * 5),
                 d_qoff = dcarve((R + 1) * 4), d_qn = dcarve(u_cap + 16);
    e = dev_alloc(&b->d_bam, dof, "device BAM chunk");
    if (e != cudaSuccess) { exlr_batch_free(b); return cuda_fail(e, "cudaMalloc(BAM chunk)"); }
    char* ds = (char*)b->d_bam;
    DevBam& D = b->db;
    D.ctrl = (BamCtrl*)(ds + d_ctrl);
    D.scan_x = (unsigned long long*)(ds + d_ctrl + sizeof(BamCtrl)); D.scan_y = D.scan_x + tiles; D.scan_z = D.scan_y + tiles;
    b->bam_zero_bytes = sizeof(BamCtrl) + (size_t)tiles * 24;
    D.comp = (const uint8_t*)(ds + d_comp); b->d_btab = (BgzfBlock*)(ds + d_tab); D.U = (uint8_t*)(ds + d_u);
    uint32_t* pb = (uint32_t*)(ds + d_blk);
    D.spec = pb; D.cnt = pb + tab_n; D.exitp = pb + 2 * tab_n; D.kind = pb + 3 * tab_n;
    D.blk_start = pb + 4 * tab_n; D.blk_base = pb + 5 * tab_n;
    D.rec_start = (uint32_t*)(ds + d_rec);
    uint32_t* pr = (uint32_t*)(ds + d_per);
    D.ncig = pr; D.salen = pr + R; D.qlen = pr + 2 * R; D.cig_src = pr + 3 * R; D.sa_src = pr + 4 * R;
    D.qname_off = (uint32_t*)(ds + d_qoff); D.qnames = (uint8_t*)(ds + d_qn);
    D.max_reads = (uint32_t)R; D.n_ref = c->n_ref;
    e = cudaHostGetDevicePointer((void**)&D.host_ctrl, b->h_bctrl, 0);
    for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&b->ev_bam[i]);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_tail_copied, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMemset(D.U, 0, u_cap + 256);
    if (e != cudaSuccess) { exlr_batch_free(b); return cuda_fail(e, "BAM chunk setup"); }
    *out = b;
    return EXLR_OK;
}

int exlr_bam_get_views(exlr_batch* b, exlr_bam_views* v)
{
    if (!b || !v || !b->is_bam) return EXLR_ERR_ARG;
    v->comp = b->h_comp; v->blocks = b->h_blocks; v->max_comp_bytes = b->max_comp; v->max_blocks = b->max_blocks; v->reserved = 0;
    v->max_tail_bytes = b->front_u;
    return EXLR_OK;
}

// Layout on the device: the chunk's own blocks inflate to U[front_u + ...); the bytes the previous chunk's walk did not consume
// (its partial last record) are copied right in front of them once that walk is done, device to device -- so a chunk's blocks
// are inflated before the previous chunk's tail is known, and nothing is inflated twice.
int exlr_bam_submit(exlr_batch* b, uint64_t comp_bytes, uint32_t n_blocks)
{
    if (!b || !b->is_bam) return EXLR_ERR_ARG;
    if (comp_bytes > b->max_comp || n_blocks > b->max_blocks) return EXLR_ERR_CAPACITY;
    CK(cudaSetDevice(b->ctx->device));
    uint64_t u = b->front_u;
    BgzfBlock* tab = b->h_btab + 1;
    for (uint32_t i = 0; i < n_blocks; i++) {                  // where every block inflates to: the prefix sum of the ISIZEs
        const exlr_bgzf_block& k = b->h_blocks[i];
        if ((uint64_t)k.comp_off + k.comp_len > comp_bytes || k.ulen > 65536u) return EXLR_ERR_BGZF;
        tab[i] = BgzfBlock{k.comp_off, k.comp_len, (uint32_t)u, k.ulen, k.crc32, {0u, 0u, 0u}};
        u += k.ulen;
    }
    cudaStream_t st = b->stream;
    DevBam& D = b->db;
    b->bam_n_new = n_blocks; b->bam_u_end = u;
    b->bam_comp_bytes = comp_bytes; b->submitted = false; b->have_timing = false;
    if (b->tail_reader) { CK(cudaStreamWaitEvent(st, b->tail_reader, 0)); b->tail_reader = nullptr; }   // the next chunk may still be copying this one's old tail
    CK(cudaEventRecord(b->ev_bam[0], st));
    CK(cudaMemsetAsync(D.ctrl, 0, b->bam_zero_bytes, st));
    if (comp_bytes) CK(cudaMemcpyAsync((void*)D.comp, b->h_comp, comp_bytes, cudaMemcpyHostToDevice, st));
    if (n_blocks) CK(cudaMemcpyAsync(b->d_btab + 1, tab, (size_t)n_blocks * sizeof(BgzfBlock), cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(b->ev_bam[1], st));
    D.blocks = b->d_btab + 1; D.n_blocks = n_blocks; D.block_index_base = 0; D.check_crc = b->ctx->bgzf_crc ? 1u : 0u;
    launch_bam_inflate(D, st);
    CK(cudaEventRecord(b->ev_bam[2], st));
    CK(cudaGetLastError());
    b->bam_state = 1;
    return EXLR_OK;
}

int exlr_bam_walk(exlr_batch* b, exlr_batch* prev, uint64_t start_off)
{
    if (!b || !b->is_bam || (prev && (!prev->is_bam || prev == b))) return EXLR_ERR_ARG;
    if (b->bam_state != 1 || (prev && prev->bam_state != 3)) return EXLR_ERR_STATE;
    CK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->stream;
    DevBam& D = b->db;
    // the previous chunk's unconsumed tail goes right in front of this chunk's own bytes (its extract has synchronised its stream)
    uint64_t tail = 0;
    if (prev) {
        const uint64_t p_end = prev->bam_u_end, p_tail = prev->bam_origin + prev->bam_info_tail;
        tail = p_end > p_tail ? p_end - p_tail : 0;
        if (tail > b->front_u) return EXLR_ERR_CAPACITY;        // a record larger than max_tail_bytes
        if (tail) {
            if (prev->ctx->device == b->ctx->device) CK(cudaMemcpyAsync(D.U + b->front_u - tail, prev->db.U + p_tail, tail, cudaMemcpyDeviceToDevice, st));
            else CK(cudaMemcpyPeerAsync(D.U + b->front_u - tail, b->ctx->device, prev->db.U + p_tail, prev->ctx->device, tail, st));
            CK(cudaEventRecord(b->ev_tail_copied, st));        // (an event of this batch's device; prev's stream can wait for it across devices)
            prev->tail_reader = b->ev_tail_copied;
        }
    }
    if (start_off > tail + (b->bam_u_end - b->front_u)) return EXLR_ERR_ARG;
    b->bam_origin = b->front_u - tail;
    // (the tail is entry 0 of the block table: a pseudo block that is already "inflated")
    const uint32_t first = tail ? 0u : 1u;
    if (tail) {
        b->h_btab[0] = BgzfBlock{0u, 0u, (uint32_t)b->bam_origin, (uint32_t)tail, 0u, {0u, 0u, 0u}};
        CK(cudaMemcpyAsync(b->d_btab, b->h_btab, sizeof(BgzfBlock), cudaMemcpyHostToDevice, st));
    }
    b->bam_n_front = tail ? 1u : 0u;
    D.blocks = b->d_btab + first; D.n_blocks = b->bam_n_front + b->bam_n_new; D.block_index_base = first;
    D.u_begin = (uint32_t)b->bam_origin; D.u_total = (uint32_t)b->bam_u_end; D.start_off = (uint32_t)(b->bam_origin + start_off);
    b->dv.hc = HostCfg{b->ctx->sms, 4, b->ctx->k1a_ctas, b->ctx->k1_waves};
    launch_bam_walk(D, b->dv, st);
    CK(cudaEventRecord(b->ev_bam[3], st));
    CK(cudaGetLastError());
    b->bam_state = 2;
    return EXLR_OK;
}

int exlr_bam_extract(exlr_batch* b, exlr_bam_info* info)
{
    if (!b || !info || !b->is_bam) return EXLR_ERR_ARG;
    if (b->bam_state != 2 && b->bam_state != 3) return EXLR_ERR_STATE;      // (3: again, after exlr_batch_grow)
    memset(info, 0, sizeof(*info));
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaStreamSynchronize(b->stream));                       // the decode header is in mapped pinned memory now
    const BamCtrl& c = *b->h_bctrl;
    info->n_blocks = b->db.n_blocks; info->u_bytes = b->bam_u_end - b->bam_origin; info->comp_bytes = b->bam_comp_bytes;
    info->bad_block = -1;
    CK(cudaEventElapsedTime(&info->h2d_ms, b->ev_bam[0], b->ev_bam[1]));
    CK(cudaEventElapsedTime(&info->inflate_ms, b->ev_bam[1], b->ev_bam[2]));
    CK(cudaEventElapsedTime(&info->walk_ms, b->ev_bam[2], b->ev_bam[3]));
    for (int i = 0; i < 4; i++) CK(cudaEventElapsedTime(&info->t_ms[i], b->ctx->ev_origin, b->ev_bam[i]));
    b->bam_state = 3;
    // (stream offsets as the caller counts them: from the first byte of the previous chunk's tail, if any, else of this chunk)
    if (c.bad_block) { info->bad_block = (int32_t)~c.bad_block; info->status = EXLR_ERR_BGZF; return info->status; }
    uint64_t n = c.n_rec;
    info->tail_off = c.tail_off - b->bam_origin; b->bam_info_tail = info->tail_off;
    if (c.capped) { info->status = EXLR_ERR_CAPACITY; return info->status; }
    if (c.bad_rec) {                                            // records before the corrupt one stand; the stream ends there
        n = (uint32_t)~c.bad_rec; info->status = EXLR_ERR_BAM_RECORD;
    } else if (c.corrupt) info->status = EXLR_ERR_BAM_RECORD;
    info->n_reads = n; info->n_ops = c.n_ops; info->n_sa_bytes = c.n_sa;
    // the records sit in the batch's device arrays exactly as exlr_upload would have left them: run the event kernels on them
    b->n_reads = n; b->dv.n_reads = (uint32_t)n; b->resident_uploaded = false; b->have_timing = false;
    if (c.bad_rec && n) {
        // (the offsets of the kept prefix are valid; its op / byte totals are read back from the device)
        unsigned long long ops = 0; uint32_t sab = 0;
        CK(cudaMemcpy(&ops, b->dv.cigar_off + n, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&sab, b->dv.sa_off + n, 4, cudaMemcpyDeviceToHost));
        info->n_ops = ops; info->n_sa_bytes = sab;
    }
    b->n_ops = info->n_ops;
    if (n == 0) { memset(b->h_ctrl, 0, sizeof(Ctrl)); b->h_line_off[0] = 0; b->submitted = true; b->formatted = true; return info->status; }
    CK(cudaEventRecord(b->ev[EV_START], b->stream));
    CK(cudaEventRecord(b->ev[EV_H2D], b->stream));
    b->h2d_bytes = b->bam_comp_bytes + (uint64_t)b->db.n_blocks * sizeof(BgzfBlock);
    const int rc = run_kernels(b, true);
    if (rc) return rc;
    b->submitted = true; b->have_timing = true;
    return info->status;
}

int exlr_bam_download_stream(exlr_batch* b, uint8_t* dst, uint64_t cap, uint64_t* n_bytes)
{
    if (!b || !n_bytes || !b->is_bam) return EXLR_ERR_ARG;
    if (b->bam_state < 1) return EXLR_ERR_STATE;
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaStreamSynchronize(b->stream));
    // state 1: the chunk's own inflated bytes; after exlr_bam_walk also the previous chunk's leftover in front of them
    const uint64_t from = b->bam_state >= 2 ? b->bam_origin : b->front_u, n = b->bam_u_end - from;
    *n_bytes = n;
    BamCtrl c;
    CK(cudaMemcpy(&c, b->db.ctrl, sizeof c, cudaMemcpyDeviceToHost));
    if (c.bad_block) return EXLR_ERR_BGZF;
    if (!dst) return EXLR_OK;                                  // (the length was all that was asked for)
    if (n > cap) return EXLR_ERR_CAPACITY;
    if (n) CK(cudaMemcpy(dst, b->db.U + from, n, cudaMemcpyDeviceToHost));
    return EXLR_OK;
}

int exlr_bam_download(exlr_batch* b, const exlr_batch_views* out, char* qnames, uint64_t qnames_cap, uint32_t* qname_off)
{
    if (!b || !out || !b->is_bam) return EXLR_ERR_ARG;
    if (b->bam_state < 2) return EXLR_ERR_STATE;
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaStreamSynchronize(b->stream));
    const BamCtrl& c = *b->h_bctrl;
    const uint64_t n = c.n_rec, ops = c.n_ops, sab = c.n_sa, qn = c.n_qn;
    if (n > out->max_reads || ops > out->max_ops || sab > out->max_sa_bytes || (qnames && qn > qnames_cap)) return EXLR_ERR_CAPACITY;
    const DevBatch& d = b->dv;
    if (ops) CK(cudaMemcpy(out->cigar, d.cigar, ops * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out->cigar_off, d.cigar_off, (n + 1) * 8, cudaMemcpyDeviceToHost));
    if (n) {
        CK(cudaMemcpy(out->pos, d.pos, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(out->tid, d.tid, n * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out->flag, d.flag, n * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(out->mapq, d.mapq, n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out->sa_kind, d.sa_kind, n, cudaMemcpyDeviceToHost));
    }
    CK(cudaMemcpy(out->sa_off, d.sa_off, (n + 1) * 4, cudaMemcpyDeviceToHost));
    if (sab) CK(cudaMemcpy(out->sa_bytes, d.sa_bytes, sab, cudaMemcpyDeviceToHost));
    if (qnames && qn) CK(cudaMemcpy(qnames, b->db.qnames, qn, cudaMemcpyDeviceToHost));
    if (qname_off) CK(cudaMemcpy(qname_off, b->db.qname_off, (n + 1) * 4, cudaMemcpyDeviceToHost));
    return EXLR_OK;
}

int exlr_get_counters(exlr_batch* b, exlr_counters* out)
{
    if (!b || !out) return EXLR_ERR_ARG;
    if (!b->submitted) return EXLR_ERR_STATE;
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaStreamSynchronize(b->stream));
    const Ctrl& c = *b->h_ctrl;
    const uint32_t room = b->dv.raw_cap - b->dv.prim_slots;
    *out = exlr_counters{c.n_flagged, c.n_short, c.n_warp, c.n_long, b->screened ? (c.n_raw < room ? c.n_raw : room) : 0ull, c.n_saev, c.n_far, c.text_bytes};
    return EXLR_OK;
}

int exlr_get_trace(exlr_batch* b, unsigned long long* out, uint32_t n_ctas)
{
    if (!b || !out || n_ctas > 8192) return EXLR_ERR_ARG;
    CK(cudaSetDevice(b->ctx->device));
    CK(cudaStreamSynchronize(b->stream));
    CK(cudaMemcpy(out, b->d_dbg, (size_t)n_ctas * 32, cudaMemcpyDeviceToHost));
    return EXLR_OK;
}

// ---- host formatter (get_alignment_event_record / get_alignment_split_record, utils.rs:196-283) ----
static inline char* put_i64(char* p, int64_t v)
{
    char tmp[24]; int n = 0;
    uint64_t u = v < 0 ? 0ull - (uint64_t)v : (uint64_t)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

static const char* const kTags[5] = {
    "excord-lr-alignment-event", "excord-lr-alignment-event-large-ins", "excord-lr-alignment-event-large-ins-one-alignments",
    "excord-lr-alignment-event-large-ins-two-alignments", "excord-lr-split-read"};

int64_t exlr_format_lines(const exlr_ctx* c, const exlr_batch* b, const exlr_result* res, uint64_t ev_begin, uint64_t ev_end,
                          int verbose, const char* qnames, const uint32_t* qname_off, char* out, uint64_t out_cap)
{
    if (!c || !b || !res || ev_begin > ev_end || ev_end > res->n_events) return EXLR_ERR_ARG;
    if (b->is_bam) return EXLR_ERR_STATE;      // the records of a BAM batch never reach the host: its lines are formatted on the device
    if (verbose && (!qnames || !qname_off)) return EXLR_ERR_ARG;
    const uint8_t* sa = b->hv.sa_bytes;
    uint64_t w = 0;
    std::vector<char> line;
    for (uint64_t i = ev_begin; i < ev_end; i++) {
        const exlr_event& e = res->events[i];
        const char* cs[2]; size_t cl[2];
        const uint32_t refs[2] = {e.lchrom, e.rchrom};
        for (int s = 0; s < 2; s++) {
            if (EXLR_CHROM_IS_SA(refs[s])) {
                const char* p = (const char*)sa + EXLR_CHROM_SA_OFF(refs[s]);
                size_t n = 0; while (p[n] != ',') n++;
                cs[s] = p; cl[s] = n;
            } else {
                const std::string& nm = c->ref_stripped[refs[s]];
                cs[s] = nm.data(); cl[s] = nm.size();
            }
        }
        const uint32_t r = e.read_idx, kind = EXLR_EV_KIND(e.meta);
        const size_t qn = verbose ? (size_t)(qname_off[r + 1] - qname_off[r]) : 0;
        line.resize(cl[0] + cl[1] + qn + 256);
        char* p = line.data();
        memcpy(p, cs[0], cl[0]); p += cl[0]; *p++ = '\t';
        p = put_i64(p, e.lstart); *p++ = '\t'; p = put_i64(p, e.lend); *p++ = '\t';
        p = put_i64(p, EXLR_EV_LSTRAND(e.meta)); *p++ = '\t';
        memcpy(p, cs[1], cl[1]); p += cl[1]; *p++ = '\t';
        p = put_i64(p, e.rstart); *p++ = '\t'; p = put_i64(p, e.rend); *p++ = '\t';
        p = put_i64(p, EXLR_EV_RSTRAND(e.meta)); *p++ = '\t';
        p = put_i64(p, (int64_t)EXLR_EV_NUM(e.meta));
        if (verbose) {
            *p++ = '\t';
            const char* tag = kTags[kind < 5 ? kind : 4]; size_t tl = strlen(tag);
            memcpy(p, tag, tl); p += tl; *p++ = '\t';
            memcpy(p, qnames + qname_off[r], qn); p += qn;
            memcpy(p, "\tstrand:", 8); p += 8;
            p = put_i64(p, (b->hv.flag[r] & 0x10) ? -1 : 1);
            memcpy(p, "\tflag:", 6); p += 6;
            p = put_i64(p, (int64_t)b->hv.flag[r]);
        }
        *p++ = '\n';
        const size_t n = (size_t)(p - line.data());
        if (w + n <= out_cap && out) memcpy(out + w, line.data(), n);
        w += n;
    }
    return (int64_t)w;
}

}  // extern "C"
