// exlr_main.cpp — `excord-lr-b200`: the reference CLI (reference src/main.rs:29-156) in front of libexlr_cuda.so.
//
//   reader thread : BGZF/BAM (-t inflate threads) -> packer -> exlr_submit, batches round-robin over the GPUs
//   writer thread : in batch order, exlr_wait_text (lines formatted on the device) or, with -v, exlr_wait + exlr_format_lines
//                   -> output file (record order)
//
// Same flags, same startup messages/exit codes, same output bytes as the reference for BAM input.  There is no CPU
// path: without a B200 the program exits with the library's error.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

#include "../../include/exlr.h"
#include "bam_chunker.hpp"
#include "bam_reader.hpp"
#include "packer.hpp"

using namespace exlr_host;

struct Cli {
    std::string bam, out, reference; bool has_reference = false;
    exlr_params p{};
    unsigned long long thread = 8;
    bool not_merge = false, debug = false, verbose = false;
    int gpus = 0;                       // 0 = all visible
    bool stats = false;                 // --stats: stage times on stderr
    unsigned long long batch_reads = 131072, batch_ops = 0, batch_sa = 0, batch_events = 0;
    bool eager_alloc = false;           // --eager-alloc: every batch of every GPU exists before the first chunk is read (steady-state timing; slower start)
    bool host_reader = false;           // --host-reader: inflate + parse BAM on host threads (zlib) even when the GPU decoder applies
    unsigned long long chunk_mb = 128, chunk_blocks = 0, max_record_mb = 64, slots = 0;   // GPU decoder: compressed bytes / BGZF blocks per chunk
};

static void usage(FILE* f)
{
    fputs("\nExtract Structural Variation Signals from Long-Read BAMs (B200 build of the signal-extraction path)\n\n"
          "Usage: excord-lr-b200 [OPTIONS] --bam <BAM> --out <OUT>\n\nOptions:\n"
          "  -b, --bam <BAM>                          Path to BAM file (BAM, or SAM text plain / BGZF; '-' or a pipe streams from stdin)\n"
          "  -r, --reference <REFERENCE>              Path to reference, used for CRAM file\n"
          "  -Q, --mapq <MAPQ>                        Minimal MapQ [default: 1]\n"
          "  -F, --exclude-flag <EXCLUDE_FLAG>        Exclude Flags [default: 1796]\n"
          "  -S, --exclude-secondary                  Exclude Secondary Alignment\n"
          "  -U, --exclude-unmapped                   Exclude Unmapped Alignment\n"
          "  -t, --thread <THREAD>                    Threads [default: 8]\n"
          "  -i, --indel-min <INDEL_MIN>              Minimal length to define an SV events in CIGAR [default: 50]\n"
          "  -m, --merge-min <MERGE_MIN>              Threshold to merge two adjacent events [default: 5]\n"
          "      --ins-clip-min <INS_CLIP_MIN>        Minimal length of hard-clip and soft-clip to define a large insertion signal [default: 1000]\n"
          "  -n, --not-merge                          Not merge (accepted and ignored, as in the reference)\n"
          "  -o, --out <OUT>                          Output file name ('-' writes to stdout)\n"
          "  -s, --split-only                         Only report split-read event\n"
          "  -p, --max-pct-overlap <MAX_PCT_OVERLAP>  Percent of overlap to discard a potential false positive record [default: 0] (alias: --pct-overlap)\n"
          "  -k, --max-supp-alignm <MAX_SUPP_ALIGNM>  Maximal number of SA to include a record [default: 4]\n"
          "  -d, --debug                              Debug\n"
          "  -v, --verbose                            Verbose output\n"
          "      --gpus <N>                           GPUs to shard batches over [default: all visible]\n"
          "      --batch-reads <N>                    Records per GPU batch [default: 131072]\n"
          "      --batch-events <N>                   Output lines a batch has room for at first [default: 4 x batch-reads + 4096]; grown on demand\n"
          "      --host-reader                        Inflate and parse BAM on host threads (-t) instead of on the GPU\n"
          "      --chunk-mb <N>                       GPU BAM decoder: compressed megabytes per chunk [default: 128]\n"
          "      --chunk-blocks <N>                   GPU BAM decoder: BGZF blocks per chunk [default: by --chunk-mb]\n"
          "      --max-record-mb <N>                  GPU BAM decoder: largest BAM record, in megabytes [default: 64]\n"
          "      --slots <N>                          Batches in flight per GPU [default: 5 with the GPU BAM decoder, else 3]\n"
          "      --eager-alloc                        Allocate every batch before the input is read (default: while the first ones work)\n"
          "      --stats                              Print stage times to stderr\n"
          "  -h, --help                               Print help\n"
          "  -V, --version                            Print version\n", f);
}

static bool parse_u(const char* s, unsigned long long max, unsigned long long* out)
{
    if (!s || !*s) return false;
    char* e = nullptr;
    if (*s == '-') return false;
    unsigned long long v = strtoull(s, &e, 10);
    if (*e || v > max) return false;
    *out = v; return true;
}

static int parse_cli(int argc, char** argv, Cli& c)
{
    exlr_params_default(&c.p);
    bool have_bam = false, have_out = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i], val; bool has_val = false;
        if (a.rfind("--", 0) == 0) { size_t eq = a.find('='); if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; } }
        else if (a.size() > 2 && a[0] == '-' ) { val = a.substr(2); a = a.substr(0, 2); has_val = true; }      // -Q1
        auto value = [&](const char** out) -> bool { if (has_val) { *out = val.c_str(); return true; } if (i + 1 < argc) { *out = argv[++i]; return true; } return false; };
        auto flagopt = [&](bool* b) -> bool { if (has_val && a.size() == 2) { fprintf(stderr, "error: unexpected value for '%s'\n", a.c_str()); return false; } *b = true; return true; };
        const char* v = nullptr; unsigned long long u = 0;
#define NEEDV() do { if (!value(&v)) { fprintf(stderr, "error: a value is required for '%s'\n", a.c_str()); return 2; } } while (0)
#define NUM(max) do { NEEDV(); if (!parse_u(v, (max), &u)) { fprintf(stderr, "error: invalid value '%s' for '%s'\n", v, a.c_str()); return 2; } } while (0)
        if (a == "-b" || a == "--bam") { NEEDV(); c.bam = v; have_bam = true; }
        else if (a == "-o" || a == "--out") { NEEDV(); c.out = v; have_out = true; }
        else if (a == "-r" || a == "--reference") { NEEDV(); c.reference = v; c.has_reference = true; }
        else if (a == "-Q" || a == "--mapq") { NUM(255); c.p.mapq = (uint8_t)u; }
        else if (a == "-F" || a == "--exclude-flag") { NUM(65535); c.p.exclude_flag = (uint16_t)u; }
        else if (a == "-S" || a == "--exclude-secondary") { bool b; if (!flagopt(&b)) return 2; c.p.exclude_secondary = 1; }
        else if (a == "-U" || a == "--exclude-unmapped") { bool b; if (!flagopt(&b)) return 2; c.p.exclude_unmapped = 1; }
        else if (a == "-t" || a == "--thread") { NUM(~0ull); c.thread = u; }
        else if (a == "-i" || a == "--indel-min") { NUM(0xffffffffull); c.p.indel_min = (uint32_t)u; }      // -i belongs to indel_min
        else if (a == "-m" || a == "--merge-min") { NUM(0xffffffffull); c.p.merge_min = (uint32_t)u; }
        else if (a == "--ins-clip-min") { NUM(0xffffffffull); c.p.ins_clip_min = (uint32_t)u; }
        else if (a == "-n" || a == "--not-merge") { if (!flagopt(&c.not_merge)) return 2; }                   // never read by the reference
        else if (a == "-s" || a == "--split-only") { bool b; if (!flagopt(&b)) return 2; c.p.split_only = 1; }
        else if (a == "-p" || a == "--max-pct-overlap" || a == "--pct-overlap") {
            NEEDV(); char* e = nullptr; double d = strtod(v, &e);
            if (e == v || *e) { fprintf(stderr, "error: invalid value '%s' for '%s'\n", v, a.c_str()); return 2; }
            c.p.max_pct_overlap = d;
        }
        else if (a == "-k" || a == "--max-supp-alignm") { NUM(~0ull); c.p.max_supp_alignm = u; }
        else if (a == "-d" || a == "--debug") { if (!flagopt(&c.debug)) return 2; }
        else if (a == "-v" || a == "--verbose") { if (!flagopt(&c.verbose)) return 2; }
        else if (a == "--gpus") { NUM(64); c.gpus = (int)u; }
        else if (a == "--batch-reads") { NUM(1ull << 30); c.batch_reads = u ? u : 1; }
        else if (a == "--batch-events") { NUM(0xfff00000ull - 1); c.batch_events = u; }
        else if (a == "--host-reader") { if (!flagopt(&c.host_reader)) return 2; }
        else if (a == "--chunk-mb") { NUM(2048); c.chunk_mb = u ? u : 1; }
        else if (a == "--chunk-blocks") { NUM(30000); c.chunk_blocks = u; }
        else if (a == "--max-record-mb") { NUM(1023); c.max_record_mb = u ? u : 1; }
        else if (a == "--slots") { NUM(16); c.slots = u; }
        else if (a == "--eager-alloc") { if (!flagopt(&c.eager_alloc)) return 2; }
        else if (a == "--stats") { if (!flagopt(&c.stats)) return 2; }
        else if (a == "-h" || a == "--help") { usage(stdout); return -1; }
        else if (a == "-V" || a == "--version") { puts("excord-LR 0.1.17"); return -1; }
        else { fprintf(stderr, "error: unexpected argument '%s' found\n\nFor more information, try '--help'.\n", argv[i]); return 2; }
    }
    if (!have_bam || !have_out) {
        fprintf(stderr, "error: the following required arguments were not provided:\n%s%s\nFor more information, try '--help'.\n",
                have_bam ? "" : "  --bam <BAM>\n", have_out ? "" : "  --out <OUT>\n");
        return 2;
    }
    return 0;
}

static bool is_file(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode); }
static bool is_dir(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
// a FIFO or character device (/dev/stdin, a named pipe, a process substitution): streamed like "-"
static bool is_stream(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && (S_ISFIFO(st.st_mode) || S_ISCHR(st.st_mode)); }

struct Slot {
    exlr_batch* b = nullptr; PackedBatch pk; int gpu = 0;
    // GPU BAM decoder: the chunk this slot holds
    exlr_bam_views bv{}; uint32_t n_new = 0; uint64_t walk_start = 0, comp_bytes = 0; exlr_bam_info info{}; int prev_slot = -1;
};

// One GPU of the run: its context and batch slots are created by its own thread, and only once a batch is headed for it
// (a small input never touches GPUs 1..n-1; a large one brings them all up side by side while GPU 0 already works).
struct Gpu {
    exlr_ctx* ctx = nullptr;
    std::deque<int> freeq;              // slots of this GPU ready to be packed (guarded by the run's mutex)
    int state = 0;                      // 0 untouched, 1 starting, 2 ready, -1 failed
    std::thread th;
};

int main(int argc, char** argv)
{
    const auto t_start = std::chrono::steady_clock::now();
    Cli cli;
    int rc = parse_cli(argc, argv, cli);
    if (rc < 0) return 0;
    if (rc) return rc;
    if (cli.debug)          // println!("{:?}", &cli)  (src/main.rs:109-111)
        printf("Cli { bam: \"%s\", reference: %s%s%s, mapq: %u, exclude_flag: %u, exclude_secondary: %s, exclude_unmapped: %s, thread: %llu, "
               "indel_min: %u, merge_min: %u, ins_clip_min: %u, not_merge: %s, out: \"%s\", split_only: %s, max_pct_overlap: %.1f, "
               "max_supp_alignm: %llu, debug: true, verbose: %s }\n",
               cli.bam.c_str(), cli.has_reference ? "Some(\"" : "None", cli.has_reference ? cli.reference.c_str() : "", cli.has_reference ? "\")" : "",
               cli.p.mapq, cli.p.exclude_flag, cli.p.exclude_secondary ? "true" : "false", cli.p.exclude_unmapped ? "true" : "false", cli.thread,
               cli.p.indel_min, cli.p.merge_min, cli.p.ins_clip_min, cli.not_merge ? "true" : "false", cli.out.c_str(),
               cli.p.split_only ? "true" : "false", cli.p.max_pct_overlap, (unsigned long long)cli.p.max_supp_alignm, cli.verbose ? "true" : "false");
    // startup checks: messages on stdout, exit 1 (src/main.rs:113-151).  Beyond the reference: "-" (and any FIFO / character
    // device) streams the input from a pipe, `-o -` writes the lines to stdout -- the reference accepts regular files only.
    const bool in_stream = cli.bam == "-" || is_stream(cli.bam), out_stdout = cli.out == "-";
    if (!in_stream && !is_file(cli.bam)) { printf("Ivalid BAM file path: %s \n", cli.bam.c_str()); return 1; }
    if (!out_stdout) {
        std::string parent;
        size_t sl = cli.out.find_last_of('/');
        parent = sl == std::string::npos ? "" : (sl == 0 ? "/" : cli.out.substr(0, sl));
        std::string abs = parent;
        if (parent.empty() || parent[0] != '/') { char cwd[4096]; if (getcwd(cwd, sizeof cwd)) abs = std::string(cwd) + (parent.empty() ? "" : "/" + parent); }
        if (!is_dir(abs)) { printf("Output directory does not exists: %s \n", abs.c_str()); return 1; }
    }
    FILE* fo = out_stdout ? stdout : fopen(cli.out.c_str(), "wb");     // created before the BAM is opened (src/main.rs:135 vs 137)
    if (!fo) { fprintf(stderr, "cannot create %s\n", cli.out.c_str()); return 101; }
    {
        std::string low = cli.bam; std::transform(low.begin(), low.end(), low.begin(), ::tolower);
        if (low.find(".cram") != std::string::npos) {
            if (!cli.has_reference) { puts("excord-lr is running on CRAM file, reference(-r) is required."); return 1; }
            fputs("CRAM input is not supported by this build (BAM only)\n", stderr); return 1;
        }
    }
    if (cli.thread == 0) { fputs("thread pool of 0 threads: the reference panics here\n", stderr); return 101; }

    // BGZF-compressed BAM in a regular file: the GPU inflates it and walks the records (exlr_bam_*), the host only hops over
    // block headers.  Everything else (SAM text, pipes, --host-reader) goes through the host reader and the packer.
    BgzfBamStream bs;
    bs.threads = (int)std::min<unsigned long long>(cli.thread, 64);
    const bool device_bam = !cli.host_reader && !in_stream && bs.open(cli.bam);
    if (!device_bam && !bs.error.empty() && !cli.host_reader && !in_stream) { fprintf(stderr, "%s\n", bs.error.c_str()); return 101; }
    BamReader rd;
    if (!device_bam && !rd.open(in_stream && cli.bam == "-" ? "-" : cli.bam, (int)std::min<unsigned long long>(cli.thread, 256))) { fprintf(stderr, "%s\n", rd.error.c_str()); return 101; }
    const std::vector<std::string>& ref_names = device_bam ? bs.ref_names : rd.ref_names;

    int ndev = exlr_device_count();
    if (ndev <= 0) { fprintf(stderr, "no usable B200: %s (%s)\n", exlr_strerror(EXLR_ERR_CUDA), exlr_last_cuda_error()); return 3; }
    if (cli.gpus > 0) ndev = std::min(ndev, cli.gpus);
    std::vector<const char*> names; for (auto& s : ref_names) names.push_back(s.c_str());
    const unsigned long long R = cli.batch_reads;
    const unsigned long long OPS = cli.batch_ops ? cli.batch_ops : std::max<unsigned long long>(R * 64, 4ull << 20);
    const unsigned long long SAB = cli.batch_sa ? cli.batch_sa : std::max<unsigned long long>(R * 64, 1ull << 20);
    const unsigned long long EVS = cli.batch_events ? cli.batch_events : (device_bam ? 0 : 4 * R + 4096);
    // GPU decoder: chunk geometry.  A chunk never needs more than the file holds; some blocks are kept free for the blocks of the
    // previous chunk that are repeated in front (from the one its last, partial record begins in)
    unsigned long long file_bytes = 0;
    { struct stat st; if (device_bam && stat(cli.bam.c_str(), &st) == 0) file_bytes = (unsigned long long)st.st_size; }
    const unsigned long long chunk_bytes = std::min<unsigned long long>(cli.chunk_mb << 20, file_bytes + 65536) + (1u << 20);
    const uint32_t chunk_blocks = (uint32_t)(cli.chunk_blocks ? cli.chunk_blocks : std::min<unsigned long long>(30000, std::max<unsigned long long>(chunk_bytes / 16384, 64)));
    const unsigned long long tail_bytes = cli.max_record_mb << 20;      // the longest leftover a chunk can inherit: one record, however many chunks it spans
    int per_gpu = cli.slots ? (int)std::max<unsigned long long>(cli.slots, 2) : (device_bam ? 5 : 3);
    if (device_bam && !cli.slots) {     // a file of few chunks never fills five slots per GPU: do not allocate (and pin) what cannot be used
        const unsigned long long est = file_bytes / std::max<unsigned long long>(cli.chunk_mb << 20, 1) + 1;
        per_gpu = (int)std::min<unsigned long long>(per_gpu, std::max<unsigned long long>(2, (est + ndev - 1) / ndev + 1));
    }
    std::vector<Slot> slots((size_t)ndev * per_gpu);
    std::vector<Gpu> gpus(ndev);

    // submitted batches travel to the writer in submission order; free slots travel back to their GPU's queue
    std::mutex mu; std::condition_variable cv;
    std::deque<int> inflight; bool done = false; int fatal = 0;

    auto start_gpu = [&](int g) {        // caller holds mu
        if (gpus[g].state != 0) return;
        gpus[g].state = 1;
        gpus[g].th = std::thread([&, g]() {
            exlr_ctx* ctx = nullptr;
            int st = exlr_create(&cli.p, g, names.data(), (int)names.size(), &ctx);
            // without -v the lines are formatted on the device (kernels 5a/5b) and the D2H copy carries the final bytes;
            // -v lines carry the read name, which stays on the host: those are formatted here
            if (!st) st = exlr_set_option(ctx, EXLR_OPT_DEVICE_FORMAT, (cli.verbose && !device_bam) ? 0 : 1);
            if (!st && device_bam) st = exlr_set_option(ctx, EXLR_OPT_VERBOSE_TEXT, cli.verbose ? 1 : 0);   // the read names are on the device there
            { std::lock_guard<std::mutex> lk(mu); gpus[g].ctx = ctx; }
            // each slot is handed to the reader as soon as it exists: the first chunk is on its way while the others' buffers
            // are still being allocated and pinned
            for (int k = 0; k < per_gpu && !st; k++) {
                Slot& s = slots[(size_t)g * per_gpu + k];
                s.gpu = g;
                if (device_bam) {
                    st = exlr_bam_batch_alloc(ctx, chunk_bytes, chunk_blocks, tail_bytes, EVS, &s.b);
                    if (!st) st = exlr_bam_get_views(s.b, &s.bv);
                } else {
                    st = exlr_batch_alloc(ctx, R, OPS, SAB, EVS, &s.b);
                    if (st) break;
                    exlr_batch_get_views(s.b, &s.pk.v);
                    s.pk.keep_qnames = cli.verbose;
                    s.pk.reset();
                }
                if (st) break;
                std::lock_guard<std::mutex> lk(mu);
                gpus[g].state = 2; gpus[g].freeq.push_back(g * per_gpu + k);
                cv.notify_all();
            }
            if (st) {
                fprintf(stderr, "GPU %d: %s (%s)\n", g, exlr_strerror(st), exlr_last_cuda_error());
                std::lock_guard<std::mutex> lk(mu);
                gpus[g].state = -1; fatal = 3;
                cv.notify_all();
            }
        });
    };
    { std::lock_guard<std::mutex> lk(mu); start_gpu(0); if (cli.eager_alloc) for (int k = 1; k < ndev; k++) start_gpu(k); }
    if (cli.eager_alloc) {               // the stream clock below then starts with every batch of every GPU in place
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { if (fatal) return true; for (auto& g : gpus) if ((int)g.freeq.size() < per_gpu) return false; return true; });
    }

    using clk = std::chrono::steady_clock;
    auto secs = [](clk::duration d) { return std::chrono::duration<double>(d).count(); };
    double t_wait = 0, t_format = 0, t_write = 0; uint64_t n_lines = 0, n_batches = 0, n_regrown = 0;
    std::thread writer([&]() {
        std::vector<char> text;
        for (;;) {
            int si;
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !inflight.empty() || done; }); if (inflight.empty()) return; si = inflight.front(); inflight.pop_front(); }
            Slot& s = slots[si];
            exlr_result res;
            auto t0 = clk::now();
            const char* dtext = nullptr; uint64_t dbytes = 0;
            bool host_format = cli.verbose && !device_bam;
            int st = 0;
            uint64_t cap_now = 0;
            for (int attempt = 0;; attempt++) {
                st = host_format ? exlr_wait(s.b, &res) : exlr_wait_text(s.b, &res, &dtext, &dbytes);
                if (st == EXLR_ERR_TEXT_CAPACITY && !device_bam) { st = exlr_wait(s.b, &res); host_format = true; }   // unusually long lines: format here
                if ((st != EXLR_ERR_CAPACITY && st != EXLR_ERR_TEXT_CAPACITY) || attempt == 6) break;
                // an event-dense batch (small -i, many split reads at a large -k): the library says how many events it needs;
                // the records are still in the slot (pinned views, or on the device for a decoded BAM chunk), so grow the event
                // buffers and run the batch again.  (Lines too long for the text buffer of a BAM chunk: the same, doubled.)
                uint64_t need = res.n_events + res.n_events / 8 + 4096;
                if (st == EXLR_ERR_TEXT_CAPACITY) need = std::max<uint64_t>(cap_now * 2, 2 * need);
                cap_now = need;
                int g = exlr_batch_grow(s.b, need);
                if (!g) { if (device_bam) { exlr_bam_info again; g = exlr_bam_extract(s.b, &again); if (g == EXLR_ERR_BAM_RECORD) g = 0; } else g = exlr_submit(s.b, s.pk.n); }
                if (g) { st = g; break; }
                n_regrown++;
            }
            t_wait += secs(clk::now() - t0); n_batches++;
            if (st != 0 && st > -10) { fprintf(stderr, "exlr_wait: %s (%s)\n", exlr_strerror(st), exlr_last_cuda_error()); std::lock_guard<std::mutex> lk(mu); fatal = 3; cv.notify_all(); return; }
            uint64_t n_ev = res.n_events;
            if (host_format) {
                if (st <= -10) n_ev = res.line_off[res.err_read] + res.n_err_lines;   // what the reference had written when it panicked
                const char* qn = cli.verbose ? s.pk.qnames.data() : nullptr;
                static const char kEmpty = 0;
                if (cli.verbose && !qn) qn = &kEmpty;
                t0 = clk::now();
                int64_t need = exlr_format_lines(gpus[s.gpu].ctx, s.b, &res, 0, n_ev, cli.verbose, qn, cli.verbose ? s.pk.qname_off.data() : nullptr, nullptr, 0);
                if (need > 0) {
                    text.resize((size_t)need);
                    exlr_format_lines(gpus[s.gpu].ctx, s.b, &res, 0, n_ev, cli.verbose, qn, cli.verbose ? s.pk.qname_off.data() : nullptr, text.data(), (uint64_t)need);
                    t_format += secs(clk::now() - t0); t0 = clk::now();
                    fwrite(text.data(), 1, (size_t)need, fo);
                    t_write += secs(clk::now() - t0);
                }
            } else if (dbytes) {                                          // (on a failing record: already cut to the lines written before the panic)
                t0 = clk::now();
                fwrite(dtext, 1, (size_t)dbytes, fo);
                t_write += secs(clk::now() - t0);
            }
            n_lines += n_ev;
            if (st <= -10) {
                fflush(fo);
                fprintf(stderr, "excord-lr-b200: record %u of a batch: %s\n", res.err_read, exlr_strerror(st));
                std::lock_guard<std::mutex> lk(mu); fatal = 101; cv.notify_all(); return;
            }
            { std::lock_guard<std::mutex> lk(mu); gpus[s.gpu].freeq.push_back(si); }
            cv.notify_all();
        }
    });

    // batch `seq` goes to GPU seq % ndev (round robin); the writer re-assembles in submission order
    auto acquire = [&](uint64_t seq) -> int {
        const int g = (int)(seq % (uint64_t)ndev);
        std::unique_lock<std::mutex> lk(mu);
        if (seq == 1) for (int k = 1; k < ndev; k++) start_gpu(k);       // more than one batch: bring the other GPUs up, all at once
        start_gpu(g);
        cv.wait(lk, [&] { return !gpus[g].freeq.empty() || fatal; });
        if (fatal) return -1;
        int si = gpus[g].freeq.front(); gpus[g].freeq.pop_front(); return si;
    };
    auto submit = [&](int si) -> bool {
        Slot& s = slots[si];
        int st = exlr_submit(s.b, s.pk.n);
        if (st) { fprintf(stderr, "exlr_submit: %s (%s)\n", exlr_strerror(st), exlr_last_cuda_error()); std::lock_guard<std::mutex> lk(mu); fatal = 3; cv.notify_all(); return false; }
        { std::lock_guard<std::mutex> lk(mu); inflight.push_back(si); }
        cv.notify_all();
        return true;
    };

    uint64_t seq = 0;
    int cur = acquire(seq);
    const auto t_begin = clk::now();
    const double t_setup = secs(t_begin - t_start);                      // CUDA context + buffers of the first GPU, BAM header
    BamRecordView r;
    uint64_t n_rec = 0;
    double t_read = 0, t_settle = 0, dev_h2d_ms = 0, dev_inflate_ms = 0, dev_walk_ms = 0; uint64_t u_bytes_total = 0, n_chunks = 0;
    if (device_bam) {
        // ---- GPU decoder: chunk k = [blocks of chunk k-1 from its partial last record on] + [new blocks]
        auto fail = [&](const char* what, int st) { fprintf(stderr, "%s: %s (%s)\n", what, exlr_strerror(st), exlr_last_cuda_error()); std::lock_guard<std::mutex> lk(mu); fatal = 3; cv.notify_all(); };
        auto hand_over = [&](int si) { { std::lock_guard<std::mutex> lk(mu); inflight.push_back(si); } cv.notify_all(); };
        // waits for chunk `si`'s record walk and starts its event kernels; false = stop reading (error, or the stream ends here)
        bool settled_ok = false;                            // the last settle() left a usable result in its slot
        auto release_slot = [&](int si) { { std::lock_guard<std::mutex> lk(mu); gpus[slots[si].gpu].freeq.push_back(si); } cv.notify_all(); };
        auto settle = [&](int si) -> bool {
            Slot& s = slots[si];
            settled_ok = false;
            auto t0 = clk::now();
            int st = exlr_bam_extract(s.b, &s.info);
            if (st == EXLR_ERR_BGZF && s.info.bad_block > 0) {
                // a block that does not inflate: like the reference's reader error, the stream ends there and the records before it stand
                const uint32_t keep = (uint32_t)s.info.bad_block;
                st = exlr_bam_submit(s.b, s.comp_bytes, keep);
                if (!st) st = exlr_bam_walk(s.b, s.prev_slot >= 0 ? slots[s.prev_slot].b : nullptr, s.walk_start);
                if (!st) st = exlr_bam_extract(s.b, &s.info);
                if (st == 0 || st == EXLR_ERR_BAM_RECORD) { s.info.status = EXLR_ERR_BGZF; st = EXLR_ERR_BGZF; }
            }
            t_settle += secs(clk::now() - t0);
            if (getenv("EXLR_BAM_TRACE")) fprintf(stderr, "chunk %llu gpu %d: start %.2f h2d-done %.2f inflate-done %.2f walk-done %.2f ms; settled at host %.2f ms\n", (unsigned long long)n_chunks, s.gpu, s.info.t_ms[0], s.info.t_ms[1], s.info.t_ms[2], s.info.t_ms[3], secs(clk::now() - t_begin) * 1e3);
            dev_h2d_ms += s.info.h2d_ms; dev_inflate_ms += s.info.inflate_ms; dev_walk_ms += s.info.walk_ms; u_bytes_total += s.info.u_bytes; n_chunks++;
            if (st != 0 && st != EXLR_ERR_BGZF && st != EXLR_ERR_BAM_RECORD) { fail("exlr_bam_extract", st); return false; }
            if (st == EXLR_ERR_BGZF && s.info.status != EXLR_ERR_BGZF) return false;
            if (st == EXLR_ERR_BGZF && s.info.bad_block >= 0) return false;         // (bad_block stays set only when nothing of this chunk is usable)
            n_rec += s.info.n_reads;
            settled_ok = true;
            return st == 0;
        };
        auto hand_over_settled = [&](int si) { if (settled_ok) hand_over(si); else release_slot(si); settled_ok = false; };
        // Two threads share the input side.  The READER reads chunks into free batches and submits them (H2D + inflate) as fast as
        // batches come back from the writer, so that several chunks' inflate kernels share the GPU(s) (one chunk alone leaves most
        // warp slots empty and DEFLATE is latency-bound per block).  This thread drives the CHAIN walk(k) -> settle(k) -> walk(k+1)
        // behind it, one chunk at a time: where chunk k+1's records begin is only known once chunk k has been walked.
        std::mutex pm; std::condition_variable pcv;
        std::deque<int> pend;                               // submitted, not yet walked, in order        (guarded by pm)
        bool input_done = false, stop_reading = false;      // reader has left; chain asks it to leave  (guarded by pm)
        auto release = [&](int si) { { std::lock_guard<std::mutex> lk(mu); gpus[slots[si].gpu].freeq.push_back(si); } cv.notify_all(); };
        const int first_slot = cur;                         // the batch acquired above, for chunk 0
        cur = -1;
        std::thread reader([&]() {
            int have = first_slot;
            for (uint64_t sq = 0;; sq++) {
                const int si = have >= 0 ? have : acquire(sq);
                have = -1;
                if (si < 0) break;                          // fatal
                { std::lock_guard<std::mutex> lk(pm); if (stop_reading) { release(si); break; } }
                Slot& s = slots[si];
                auto t0 = clk::now();
                size_t new_bytes = 0;
                const size_t nb = bs.read_blocks(s.bv.comp, (size_t)s.bv.max_comp_bytes, s.bv.blocks, s.bv.max_blocks, &new_bytes);
                t_read += secs(clk::now() - t0);
                if (nb == 0) { release(si); break; }
                s.comp_bytes = new_bytes; s.n_new = (uint32_t)nb;
                const int st = exlr_bam_submit(s.b, new_bytes, (uint32_t)nb);
                if (st) { fail("exlr_bam_submit", st); release(si); break; }
                { std::lock_guard<std::mutex> lk(pm); pend.push_back(si); }
                pcv.notify_all();
            }
            { std::lock_guard<std::mutex> lk(pm); input_done = true; }
            pcv.notify_all();
        });
        int held = -1;                                      // settled, its leftover not yet handed to the next chunk
        uint64_t start = bs.first_record_off;
        bool stopped = false;                               // the stream ends here: what the reader still submits is dropped
        for (;;) {
            int si;
            {
                std::unique_lock<std::mutex> lk(pm);
                pcv.wait(lk, [&] { return !pend.empty() || input_done; });
                if (pend.empty()) break;
                si = pend.front(); pend.pop_front();
            }
            if (stopped || fatal) { release(si); continue; }
            Slot& s = slots[si];
            s.walk_start = start; s.prev_slot = held;
            const int st = exlr_bam_walk(s.b, held >= 0 ? slots[held].b : nullptr, start);
            // the previous chunk's leftover (a partial last record) is copied in front of this one on the device: its batch goes to
            // the writer only now that the copy is in this chunk's stream (the library orders the batch's reuse behind it)
            if (held >= 0) { hand_over_settled(held); held = -1; }
            if (st == EXLR_ERR_CAPACITY) fprintf(stderr, "a BAM record is larger than --max-record-mb (%llu MB)\n", cli.max_record_mb);
            if (st) { fail("exlr_bam_walk", st); release(si); stopped = true; }
            else {
                const bool go_on = settle(si);
                start = 0;
                held = si;
                if (!go_on) { hand_over_settled(held); held = -1; stopped = true; }
            }
            if (stopped) { std::lock_guard<std::mutex> lk(pm); stop_reading = true; }
        }
        if (held >= 0) hand_over_settled(held);
        reader.join();
        // (a partial record left at the very end of the input is a truncated file: dropped, like a read error in the reference)
        cur = -1;
    }
    while (cur >= 0 && rd.next(r)) {
        n_rec++;
        PackedBatch* pk = &slots[cur].pk;
        if (!pk->fits(r)) {
            if (pk->n == 0 || !pk->can_ever_fit(r)) { fprintf(stderr, "record %llu does not fit a batch (CIGAR ops %u, SA bytes %u): raise --batch-reads\n", (unsigned long long)n_rec, r.n_cigar, r.sa_len); std::lock_guard<std::mutex> lk(mu); fatal = 3; break; }
            if (!submit(cur)) break;
            cur = acquire(++seq);
            if (cur < 0) break;
            pk = &slots[cur].pk; pk->reset();
        }
        pk->push(r);
    }
    if (cur >= 0 && !fatal && slots[cur].pk.n) submit(cur);
    { std::lock_guard<std::mutex> lk(mu); done = true; }
    cv.notify_all();
    writer.join();
    for (auto& g : gpus) if (g.th.joinable()) g.th.join();
    if (out_stdout) fflush(fo); else fclose(fo);
    int n_used = 0; for (auto& g : gpus) n_used += g.state == 2;
    if (cli.stats && device_bam)
        fprintf(stderr, "excord-lr-b200 GPU BAM decoder: %llu chunks, %.1f MB of BGZF -> %.1f MB of BAM; reader thread: reading the file into pinned memory %.3f s, "
                "waiting for a chunk's record walk %.3f s; device: H2D %.1f ms, inflate %.1f ms (%.1f GB/s of BAM out), record walk + gather %.1f ms\n",
                (unsigned long long)n_chunks, bs.bytes_read / 1e6, u_bytes_total / 1e6, t_read, t_settle, dev_h2d_ms, dev_inflate_ms,
                dev_inflate_ms > 0 ? u_bytes_total / 1e6 / dev_inflate_ms : 0.0, dev_walk_ms);
    if (cli.stats)
        fprintf(stderr, "excord-lr-b200: %llu records, %llu lines, %llu batches (%llu re-run with larger event buffers) on %d of %d GPU(s); setup (CUDA context, "
                "pinned + device buffers of the first GPU, BAM header) %.3f s; "
                "stream %.3f s (writer thread: waiting for the GPU %.3f s, formatting %.3f s, writing %.3f s; the reader thread %s "
                "for the whole stream)\n",
                (unsigned long long)n_rec, (unsigned long long)n_lines, (unsigned long long)n_batches, (unsigned long long)n_regrown, n_used, ndev, t_setup,
                secs(clk::now() - t_begin), t_wait, t_format, t_write, device_bam ? "reads the file and drives the GPU decoder" : "inflates, parses and packs");
    for (auto& s : slots) if (s.b) exlr_batch_free(s.b);
    for (auto& g : gpus) if (g.ctx) exlr_destroy(g.ctx);
    return fatal;
}
