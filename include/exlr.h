/*
 * exlr.h — C ABI of libexlr_cuda.so, the B200-native (sm_100a) replacement for the
 * signal-extraction hot path of excord-lr.
 *
 * The reference has no FFI seam: its hot path is the body of the per-record loop in
 * `main()` (reference src/main.rs:158-770).  This header is the seam a host reader
 * (Rust over `extern "C"`, C++, or Python/ctypes) binds instead of that loop body:
 *
 *   reference (src/main.rs)                      this ABI
 *   -------------------------------------------  -------------------------------------------
 *   Cli fields used in the loop (:47-96)         exlr_params
 *   header tid -> contig name (:198)             exlr_create(ref_names, n_ref)
 *   one `Record` per iteration (:156-158)        exlr_batch_* : a structure-of-arrays batch of records
 *   loop body :169-768                           exlr_submit / exlr_wait  (kernels 1-4 on the GPU)
 *   f.write(line) sites :395,420,448,484,515,766 exlr_result.events (ordered) + exlr_format_lines
 *   utils.rs:196-283 formatters                  exlr_format_lines
 *   unwrap()/index panics (utils.rs:109,122-135) negative status codes + exlr_result.err_read
 *
 * All entry points are plain C: pointers and sizes only, no C++/torch types, no exceptions
 * cross the boundary.  There is NO CPU fallback: every compute entry point fails with
 * EXLR_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef EXLR_H
#define EXLR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EXLR_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------------------- */
#define EXLR_OK                 0
#define EXLR_ERR_ARG           -1   /* bad argument / NULL pointer                                   */
#define EXLR_ERR_CUDA          -2   /* CUDA runtime error or no usable device (exlr_last_cuda_error) */
#define EXLR_ERR_NOMEM         -3   /* host or device allocation failed                              */
#define EXLR_ERR_CAPACITY      -4   /* batch larger than its allocation, or event buffer overflow    */
#define EXLR_ERR_STATE         -5   /* call sequence error (wait without submit, ...)                */
#define EXLR_ERR_BGZF          -7   /* exlr_bam_*: corrupt BGZF block (the reference's reader returns an error there and the loop stops quietly, main.rs:165-168) */
#define EXLR_ERR_BAM_RECORD    -8   /* exlr_bam_*: corrupt BAM record in the inflated stream (same: the records before it stand)        */
#define EXLR_ERR_TEXT_CAPACITY -6   /* exlr_wait_text: the lines need more than the batch's text buffer; format on the host */
/* Per-record conditions on which the reference panics (exit 101).  exlr_result.err_read holds the
 * smallest offending record index of the batch; the lines the reference had written by then are valid:
 * events [0, line_off[err_read] + n_err_lines). */
#define EXLR_ERR_TID          -10   /* kept record with tid<0 or tid>=n_ref (main.rs:198, contig() panics)      */
#define EXLR_ERR_CIGAR_OP     -11   /* BAM CIGAR op code > 8 (rust-htslib Cigar decode panics; main.rs:243,525) */
#define EXLR_ERR_SA_FIELDS    -12   /* SA piece with < 6 ','-separated fields (utils.rs:121-130 index panic)    */
#define EXLR_ERR_SA_POS       -13   /* SA pos not an i64 (utils.rs:122)                                         */
#define EXLR_ERR_SA_STRAND    -14   /* SA strand not "+" / "-" (utils.rs:123-127,135)                           */
#define EXLR_ERR_SA_CIGAR     -15   /* SA CIGAR: op with empty/overflowing count (utils.rs:109) or a byte outside
                                       [0-9MIDNSHP=X] (outside the supported domain)             */
#define EXLR_ERR_SA_MAPQ      -16   /* SA mapq not a u8 (utils.rs:129)                                          */
#define EXLR_ERR_SA_NM        -17   /* SA NM not an i64 (utils.rs:130)                                          */
#define EXLR_ERR_MERGE_DOMAIN -20   /* the >2-event merge loop indexes out of bounds (main.rs:664-671): the reference
                                       panics AFTER the SA arm of the same record wrote its lines (n_err_lines).  Where the loop
                                       merges or duplicates events without panicking, the library emits what the reference prints */
#define EXLR_ERR_SPLIT_COUNT  -21   /* more than 2^24 segments in one record (meta field overflow)              */

/* ---- parameters: the Cli fields read inside the loop (main.rs:47-96) ------------------ */
typedef struct exlr_params {
    uint8_t  mapq;               /* -Q  (main.rs:47-48, used :179)                  default 1    */
    uint8_t  exclude_secondary;  /* -S  (main.rs:55-56, used :169)                  default 0    */
    uint8_t  exclude_unmapped;   /* -U  (main.rs:59-60, used :174)                  default 0    */
    uint8_t  split_only;         /* -s  (main.rs:87-88, used :523)                  default 0    */
    uint16_t exclude_flag;       /* -F  (main.rs:51-52, used :185)                  default 1796 */
    uint16_t reserved0;
    uint32_t indel_min;          /* -i  (main.rs:67-68, used :553,569)              default 50   */
    uint32_t merge_min;          /* -m  (main.rs:71-72, used :615,673-700)          default 5    */
    uint32_t ins_clip_min;       /* --ins-clip-min (main.rs:75-76, used :347,353,461) default 1000 */
    uint32_t reserved1;
    double   max_pct_overlap;    /* -p  (main.rs:91-92, used :351)                  default 0.0  */
    uint64_t max_supp_alignm;    /* -k  (main.rs:95-96, used :311)                  default 4    */
} exlr_params;

/* Fills *p with the reference defaults (main.rs:47-96). */
void exlr_params_default(exlr_params* p);

/* ---- one output line ---------------------------------------------------------------- */
/* kind: which f.write site produced the line (selects the -v tag, utils.rs:219,263) */
#define EXLR_KIND_INDEL        0u  /* main.rs:766  "excord-lr-alignment-event"                        */
#define EXLR_KIND_INS_ONE_SEG  1u  /* main.rs:484  "excord-lr-alignment-event-large-ins"               */
#define EXLR_KIND_INS_ONE_ALN  2u  /* main.rs:448  "excord-lr-alignment-event-large-ins-one-alignments" */
#define EXLR_KIND_INS_TWO_ALN  3u  /* main.rs:395,420 "excord-lr-alignment-event-large-ins-two-alignments" */
#define EXLR_KIND_SPLIT        4u  /* main.rs:515  "excord-lr-split-read"                              */

/* chrom reference: bit31 clear -> tid (header name, "chr" stripped);
 * bit31 set -> byte offset into the batch's sa_bytes where the (already "chr"-stripped)
 * SA contig name starts; it ends at the next ','. */
#define EXLR_CHROM_IS_SA(c)   (((c) >> 31) & 1u)
#define EXLR_CHROM_SA_OFF(c)  ((c) & 0x7fffffffu)

/* meta: events_num (column 9) | kind | strand signs */
#define EXLR_EV_NUM(m)      ((m) & 0x00ffffffu)
#define EXLR_EV_KIND(m)     (((m) >> 24) & 0xfu)
#define EXLR_EV_LSTRAND(m)  ((((m) >> 28) & 1u) ? -1 : 1)
#define EXLR_EV_RSTRAND(m)  ((((m) >> 29) & 1u) ? -1 : 1)
#define EXLR_EV_META(num, kind, lneg, rneg) \
    (((uint32_t)(num) & 0x00ffffffu) | ((uint32_t)(kind) << 24) | ((uint32_t)((lneg) ? 1 : 0) << 28) | ((uint32_t)((rneg) ? 1 : 0) << 29))

/* 48 bytes.  Coordinates hold the exact value the reference prints: u32-wrapped (zero
 * extended) for kinds 0-3 (aligments_event.rs:16-21, main.rs:375-380), signed i64 for
 * kind 4 (split_read_event.rs:5-6). */
typedef struct exlr_event {
    int64_t  lstart, lend, rstart, rend;
    uint32_t read_idx;   /* record index inside the batch */
    uint32_t lchrom;     /* chrom reference, see above    */
    uint32_t rchrom;
    uint32_t meta;
} exlr_event;

/* ---- batch: structure-of-arrays views in pinned host memory ------------------------- */
#define EXLR_SA_NONE   0u  /* record.aux("SA") is Err            (main.rs:518)                      */
#define EXLR_SA_STRING 1u  /* Ok(Aux::String)                     (main.rs:308)                      */
#define EXLR_SA_OTHER  2u  /* Ok(other aux type): SA arm runs with the record as the only segment    */

typedef struct exlr_batch_views {
    uint32_t* cigar;      /* [max_ops]  BAM encoding len<<4|op, records back to back (no padding)  */
    uint64_t* cigar_off;  /* [max_reads+1] op offset of each record; cigar_off[0] must be 0        */
    int32_t*  pos;        /* [max_reads] 0-based leftmost position (record.pos(), main.rs:199)     */
    int32_t*  tid;        /* [max_reads] reference id (record.tid())                               */
    uint16_t* flag;       /* [max_reads] record.flags()                                            */
    uint8_t*  mapq;       /* [max_reads] record.mapq()                                             */
    uint8_t*  sa_kind;    /* [max_reads] EXLR_SA_*                                                 */
    uint32_t* sa_off;     /* [max_reads+1] byte offset of each record's SA string; sa_off[0] = 0   */
    uint8_t*  sa_bytes;   /* [max_sa_bytes] SA Z-strings back to back, no NUL                      */
    uint64_t  max_reads, max_ops, max_sa_bytes, max_events;
} exlr_batch_views;

typedef struct exlr_result {
    int32_t  status;          /* EXLR_OK or the (negative) code of the first failing record         */
    uint32_t err_read;        /* smallest failing record index (valid when status <= -10)           */
    uint64_t n_reads;
    uint64_t n_events;        /* number of output lines (on EXLR_ERR_CAPACITY: the max_events needed) */
    const exlr_event* events; /* [n_events] in reference output order               */
    const uint32_t* line_off; /* [n_reads+1] events of record i are [line_off[i], line_off[i+1])    */
    uint64_t n_kept;          /* records that passed the filters (main.rs:169-190)                  */
    uint64_t n_sa_reads;      /* kept records with an SA aux                                        */
    uint64_t n_cap_dropped;   /* kept records skipped by the -k cap (main.rs:311-313)               */
    uint64_t n_ops;           /* CIGAR ops in the batch                                             */
    uint64_t n_err_lines;     /* status <= -10: lines of record err_read itself that the reference had written before it
                                 panicked (its SA-arm lines when the indel arm panics, main.rs:395-515 vs :523-742), else 0 */
} exlr_result;

/* per-stage device times of the last submit, from CUDA events on the batch's stream (ms) */
typedef struct exlr_timing {
    float h2d_ms;       /* host->device copies of the SoA sections              */
    float classify_ms;  /* kernel 0: filter + ordered list of SA records        */
    float cigar_ms;     /* kernel 1: CIGAR scan / indel events                  */
    float sa_cigar_ms;  /* kernel 3a: per-op-type sums of SA records' CIGARs    */
    float sa_parse_ms;  /* kernel 3b: SA parse, sort, large-INS + split events  */
    float scan_ms;      /* kernel 4a: per-record line counts -> offsets         */
    float place_ms;     /* kernel 4b: ordered compaction into the event buffer  */
    float kernels_ms;   /* first kernel start -> last kernel end (the last kernel stores the result header in pinned host memory) */
    float d2h_ms;       /* what the stream still does after the last kernel (nothing but the closing event since the header is a kernel) */
    uint32_t launches;  /* kernels launched by this submit                      */
    float screen_ms;    /* kernel 1a (event screen in front of kernel 1), part of cigar_ms; 0 when it did not run */
    uint64_t h2d_bytes; /* bytes the last exlr_submit copied to the device                                        */
    uint64_t d2h_bytes; /* bytes the last submit + wait copied back (result header, and the events / line offsets /
                           lines copied behind the kernels: sized by a guess, so a little more than the result needs) */
} exlr_timing;

typedef struct exlr_ctx exlr_ctx;
typedef struct exlr_batch exlr_batch;

/* options for exlr_set_option */
#define EXLR_OPT_CIGAR_KERNEL 1  /* kernel 1 strategy.  0 = auto (default): a streaming event screen over the CIGAR array (kernel 1a), then only the
                                    records around a candidate are scanned (1b: one thread per short record; long ones by 1d from
                                    the per-step sums of 1a, see EXLR_OPT_LONG_RECORDS); event-dense batches fall back to 2.
                                    Forced: 1 = warp per record; 2 = flat TMA-staged block scan of everything; 3 = the screened path */
#define EXLR_OPT_READS_PER_CTA 2 /* 0 = auto */
#define EXLR_OPT_OVERLAP 3       /* 1 (default) = kernel 1 runs on a second stream beside kernels 0/3a/3b; 2 = the same, started once kernel 0 is done; 0 = one stream */
#define EXLR_OPT_K1_CTAS_PER_SM 4 /* 1..4 CTAs of kernel 1 per SM; 0 (default) = 3 when overlapping, else 4 */
#define EXLR_OPT_TRACE 7          /* debug: %globaltimer traces, read back with exlr_get_trace.  1 = kernel 1 / 1b per CTA / step, 2..6 = k0, k3a,
                                     k3b, k4a, k4b per CTA {start, mid, end}, 7 = timeline: entry k = {~first start, last end, -, CTAs} of kernel k */
#define EXLR_OPT_STAGE_TIMING 6   /* 1 (default) = CUDA events between the kernels, so exlr_get_timing has per-stage times */
#define EXLR_OPT_K1_WAVES 5       /* kernel 1 grid = SMs x CTAs/SM x waves (default 3) */
#define EXLR_OPT_LONG_RECORDS 10   /* screened CIGAR path, records of more than 256 ops: 0 = auto (by the batch's mean CIGAR length),
                                     1 = a warp of kernel 1b each, 2 = kernel 1c (flat block scan of the listed records),
                                     3 = kernel 1d (from the per-step sums of kernel 1a, no rescan) */
#define EXLR_OPT_K1A_CTAS_PER_SM 9 /* resident CTAs of kernel 1a per SM, 1..8 (default 8) */
#define EXLR_OPT_DEVICE_FORMAT 8  /* 1 = batches allocated from now on also format the non-verbose output lines on the device
                                     (kernels 5a/5b, utils.rs:225-236, 269-280); read them with exlr_wait_text */

#define EXLR_OPT_VERBOSE_TEXT 11  /* 1 = BAM batches (exlr_bam_batch_alloc) allocated from now on format the -v columns on the device too
                                     (tag, read name, strand, flag: utils.rs:205-223, 252-267) -- there the read names are on the device */

#define EXLR_OPT_K3_FOLD 12       /* 1 = in batches of short CIGARs kernel 3b does kernel 3a's work itself (one launch less; measured slower than the
                                     two kernels -- 53 vs 17 + 31 us on configs[1] -- so the default is 0: kernel 3a always runs) */

#define EXLR_OPT_GRAPH 13         /* 1 (default) = a batch submitted with the same shape again and again runs its kernels as one CUDA graph launch */

#define EXLR_OPT_WC_INPUT 14      /* 1 = batches allocated from now on get write-combined pinned input views (A/B for multi-GPU H2D; default 0) */

#define EXLR_OPT_K0_WALK 15       /* 1 = in batches of short CIGARs kernel 0 walks the CIGARs of the SA records it lists (kernel 3a's work) -- measured
                                     slower (SA records cluster, so a few threads walk a dozen CIGARs one after the other: 45 vs 23 + 17 us);
                                     default 0: kernel 3a always runs */

#define EXLR_OPT_BGZF_CRC 16      /* 1 (default) = exlr_bam_submit verifies the CRC-32 of every BGZF block on the device, like htslib's bgzf reader:
                                     a mismatch is EXLR_ERR_BGZF with bad_block set */

/* ---- lifecycle ---------------------------------------------------------------------- */
int  exlr_abi_version(void);
/* Number of CUDA devices with compute capability 10.x; <0 on CUDA error. */
int  exlr_device_count(void);

/* ref_names[i] is the header name of tid i (un-stripped; the library strips one leading
 * "chr" as aligments_event.rs:38-42 / split_read_event.rs:30-34 do). */
int  exlr_create(const exlr_params* p, int device, const char* const* ref_names, int n_ref, exlr_ctx** out);
void exlr_destroy(exlr_ctx* ctx);
int  exlr_set_option(exlr_ctx* ctx, int option, int64_t value);

/* Allocates pinned host staging + device buffers for one batch (own CUDA stream). */
int  exlr_batch_alloc(exlr_ctx* ctx, uint64_t max_reads, uint64_t max_ops, uint64_t max_sa_bytes,
                      uint64_t max_events, exlr_batch** out);
void exlr_batch_free(exlr_batch* b);
int  exlr_batch_get_views(exlr_batch* b, exlr_batch_views* v);
/* Replaces the batch's event / text buffers by larger ones (max_events entries); the packed records in the pinned views stay.
 * The answer to EXLR_ERR_CAPACITY: grow to exlr_result.n_events (or more) and submit the same records again.  The batch must
 * not be in flight. */
int  exlr_batch_grow(exlr_batch* b, uint64_t max_events);

/* Asynchronous: H2D of the first n_reads records of the pinned views, kernels, result
 * header D2H, all on the batch's stream.  The views must not be modified until exlr_wait. */
int  exlr_submit(exlr_batch* b, uint64_t n_reads);
/* Kernel-only variants for device-resident measurement: exlr_upload copies the views to
 * the device (synchronous); exlr_submit_resident runs the kernels on the resident copy. */
int  exlr_upload(exlr_batch* b, uint64_t n_reads);
int  exlr_submit_resident(exlr_batch* b);
/* Blocks until the batch is done and fills *res with the events + line offsets in pinned host memory (pointers valid until
 * the next submit/grow/free of this batch).  exlr_submit already copied them back behind the kernels, sized by the previous
 * batch's count, so this is normally ONE stream synchronisation; only what that guess missed is fetched afterwards.
 * Returns res->status. */
int  exlr_wait(exlr_batch* b, exlr_result* res);
/* With EXLR_OPT_DEVICE_FORMAT: blocks until the batch is done and copies the formatted lines (no -v columns) to pinned host
 * memory: *text / *n_bytes are exactly the bytes the reference writes for this batch (for a record the reference panics on:
 * the lines of the records before it).  res->events / res->line_off are NOT fetched.  EXLR_ERR_TEXT_CAPACITY: the lines do
 * not fit the text buffer (96 bytes per max_events entry); exlr_wait + exlr_format_lines still work for the batch. */
int  exlr_wait_text(exlr_batch* b, exlr_result* res, const char** text, uint64_t* n_bytes);
/* Like exlr_wait but leaves events on the device (only the result header is read). */
int  exlr_wait_resident(exlr_batch* b, exlr_result* res);
int  exlr_get_timing(exlr_batch* b, exlr_timing* t);
/* Device-side work counters of the last waited batch (diagnostics; what the per-kernel roofline figures of bench.py are
 * computed from). */
typedef struct exlr_counters {
    uint64_t flagged_steps;   /* 512-op steps of the CIGAR stream in which kernel 1a found an event candidate            */
    uint64_t claimed_short, claimed_warp, claimed_long;   /* records kernel 1b claimed, by the kernel that scans them     */
    uint64_t raw_events;      /* indel events before the merge (main.rs:549-600)                                         */
    uint64_t sa_events;       /* lines of the SA arm (main.rs:340-516)                                                   */
    uint64_t far_records;     /* records whose >2-event merge loop changed the event list (main.rs:636-742)              */
    uint64_t text_bytes;      /* bytes of the device-formatted lines (EXLR_OPT_DEVICE_FORMAT)                            */
} exlr_counters;
int  exlr_get_counters(exlr_batch* b, exlr_counters* c);
/* debug (EXLR_OPT_TRACE): 4 x u64 per CTA of kernel 1 {globaltimer at start, at first data, at end, tiles | scanned tiles << 32} */
int  exlr_get_trace(exlr_batch* b, unsigned long long* out, uint32_t n_ctas);

/* ---- BAM input decoded on the device ---------------------------------------------------------
 * Replaces, for BGZF-compressed BAM, what htslib does for the reference between the file and the record loop
 * (bam::Reader::from_path, set_threads, bam.read(&mut record): main.rs:137,155,158): the host only hops over the BGZF block
 * headers of a chunk of the file (gzip header with the BC extra field, 18 bytes; ISIZE in the last 4 bytes of the block) and
 * hands the compressed bytes over; DEFLATE, the BAM record walk (with htslib's CG:B,I long-CIGAR restore), the SA aux
 * lookup and the packing into the structure-of-arrays batch all run as kernels, and the event kernels follow on the same
 * device arrays.  A chunk is a run of whole BGZF blocks; records may straddle chunks:
 *
 *   exlr_bam_submit(b, bytes, n_blocks)   async: H2D of the chunk's blocks + inflate (needs nothing from the previous chunk, so
 *                                         the inflate kernels of consecutive chunks share the device)
 *   exlr_bam_walk(b, prev, start_off)     async: the bytes the previous chunk's walk left over (its partial last record; prev must
 *                                         have been through exlr_bam_extract; NULL for the first chunk) are copied in front of this
 *                                         chunk's, device to device; then the records from uncompressed offset start_off on,
 *                                         counted from the first of those bytes (first chunk: the end of the BAM header; else 0)
 *   exlr_bam_extract(b, &info)            waits for the walk, reports the chunk (records; tail_off = first byte not consumed, a
 *                                         partial record or the chunk's end, counted like start_off) and starts the event kernels
 *   exlr_wait_text(b, ...)                the lines, exactly as for exlr_submit with EXLR_OPT_DEVICE_FORMAT
 */
typedef struct exlr_bgzf_block {
    uint32_t comp_off;    /* offset of the block's DEFLATE data (after the gzip header + extra field) in the chunk */
    uint32_t comp_len;    /* its length: BSIZE + 1 - XLEN - 20                                                     */
    uint32_t ulen;        /* ISIZE: bytes the block inflates to (<= 65536)                                         */
    uint32_t crc32;       /* CRC32 field of the block's footer: verified on the device (EXLR_OPT_BGZF_CRC, default 1)  */
} exlr_bgzf_block;

typedef struct exlr_bam_views {
    uint8_t* comp;             /* [max_comp_bytes] pinned: the caller reads the file straight into it */
    exlr_bgzf_block* blocks;   /* [max_blocks] pinned                                                 */
    uint64_t max_comp_bytes;
    uint32_t max_blocks, reserved;
    uint64_t max_tail_bytes;   /* the longest leftover of a previous chunk (= the largest record) the batch has room for */
} exlr_bam_views;

typedef struct exlr_bam_info {
    int32_t  status;           /* EXLR_OK, EXLR_ERR_BGZF (bad_block set), EXLR_ERR_BAM_RECORD (n_reads = the records before it) */
    int32_t  bad_block;        /* first block whose DEFLATE stream is corrupt, else -1                                        */
    uint32_t n_blocks, reserved;
    uint64_t n_reads, n_ops, n_sa_bytes;
    uint64_t tail_off;         /* uncompressed offset inside this chunk of the first byte the walk did not consume            */
    uint64_t u_bytes, comp_bytes;
    float    h2d_ms, inflate_ms, walk_ms;   /* device times of the decode stages (CUDA events on the batch's stream)          */
    float    reserved2;
    float    t_ms[4];          /* the same events on the context's clock (ms since exlr_create): chunk start, H2D done, inflate done,
                                  walk + gather done -- shows how the chunks of a stream overlap on the device                */
} exlr_bam_info;

/* A batch for chunks of at most max_blocks BGZF blocks / max_comp_bytes compressed bytes, plus max_tail_bytes left over from
 * the previous chunk; its record capacity is the most such a chunk can hold, so no chunk overflows it.  Lines are always
 * formatted on the device (exlr_wait_text).  Such a batch keeps no pinned copies of the events and line offsets -- they are sized
 * by the worst-case record count -- until exlr_wait is first called on it (which then allocates them); its pinned text buffer
 * starts at 16 MB and grows when a chunk's lines need more. */
int  exlr_bam_batch_alloc(exlr_ctx* ctx, uint64_t max_comp_bytes, uint32_t max_blocks, uint64_t max_tail_bytes, uint64_t max_events, exlr_batch** out);
int  exlr_bam_get_views(exlr_batch* b, exlr_bam_views* v);
int  exlr_bam_submit(exlr_batch* b, uint64_t comp_bytes, uint32_t n_blocks);
int  exlr_bam_walk(exlr_batch* b, exlr_batch* prev, uint64_t start_off);
int  exlr_bam_extract(exlr_batch* b, exlr_bam_info* info);
/* Inspection / tests: copies the decoded structure-of-arrays batch of the last walked chunk into caller memory (`out` holds
 * caller-allocated arrays and their capacities; qnames / qname_off may be NULL). */
int  exlr_bam_download(exlr_batch* b, const exlr_batch_views* out, char* qnames, uint64_t qnames_cap, uint32_t* qname_off);
/* Inspection / tests: the inflated byte stream of the chunk (after exlr_bam_walk: with the previous chunk's leftover in front).
 * *n_bytes = its length; EXLR_ERR_BGZF if a block did not inflate; dst may be NULL to ask for the length only. */
int  exlr_bam_download_stream(exlr_batch* b, uint8_t* dst, uint64_t cap, uint64_t* n_bytes);

/* ---- host formatter: utils.rs:196-283 ------------------------------------------------- */
/* Writes the lines of events [ev_begin, ev_end) of `res` into out (capacity out_cap) and
 * returns the number of bytes the lines need (> out_cap means nothing useful was written).
 * verbose != 0 adds "\t{tag}\t{qname}\tstrand:{s}\tflag:{f}" and needs qnames/qname_off
 * (qname of record i = qnames[qname_off[i] .. qname_off[i+1])). */
int64_t exlr_format_lines(const exlr_ctx* ctx, const exlr_batch* b, const exlr_result* res,
                          uint64_t ev_begin, uint64_t ev_end, int verbose,
                          const char* qnames, const uint32_t* qname_off,
                          char* out, uint64_t out_cap);

const char* exlr_strerror(int status);
const char* exlr_last_cuda_error(void);

#ifdef __cplusplus
}
#endif
#endif /* EXLR_H */
