// exlr_device.cuh — device-side data layout shared by the kernels and the host ABI layer.
//
// Everything here restates, for the GPU, the per-record loop body of excord-lr
// (reference src/main.rs:158-770).  See DESIGN.md for the kernel map.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/exlr.h"

namespace exlr {

// ---- device error word -----------------------------------------------------------------
// key = (read_idx << 8) | rank; the smallest key wins: the smallest read, and for one read the
// condition the reference would hit first (order of main.rs:198, :243, :311-320, :523-742).
// Ctrl.err_key stores ~key (0 = no error) so one zeroing memset resets the whole block.
enum ErrRank : uint32_t {
    RANK_TID = 1, RANK_CIGAR_OP = 2, RANK_SA_FIELDS = 3, RANK_SA_POS = 4, RANK_SA_STRAND = 5, RANK_SA_CIGAR = 6,
    RANK_SA_MAPQ = 7, RANK_SA_NM = 8, RANK_MERGE_DOMAIN = 9, RANK_SPLIT_COUNT = 10
};

// ---- result header / control block (one per batch, zeroed by a memset node per submit) --
struct Ctrl {
    unsigned long long err_key;     // ~key of the winning error, 0 if none
    uint32_t n_raw;                 // raw indel events emitted by kernel 1
    uint32_t n_saev;                // SA-derived events emitted by kernel 3b
    uint32_t n_sa;                  // kept records with an SA aux (length of sa_list)
    uint32_t n_kept;
    uint32_t n_dropped;             // -k cap drops
    uint32_t n_events;              // total output lines
    uint32_t overflow;              // 1 = raw / sa / final event buffer too small
    uint32_t ticket_a, ticket_b;    // dynamic tile ids of the two chained scans
    uint32_t seg_pool_used;
    uint32_t n_flagged;             // 512-op steps of the CIGAR stream kernel 1a found an event candidate in (length of step_list)
    uint32_t n_long;                // records kernel 1b passed on to kernel 1c (length of long_list)
    uint32_t n_short;               // records kernel 1b walks itself (length of short_list)
    uint32_t ticket_c;              // dynamic tile ids of kernel 5a's chained scan
    uint32_t text_bytes;            // kernel 5a: bytes of the formatted lines
    uint32_t n_warp;                // records kernel 1b scans with a warp each (length of warp_list)
    uint32_t n_far;                 // records whose >2-event merge loop (main.rs:636-742) changed the event list (length of far_list)
    uint32_t err_lines;             // lines the failing record had written before it panicked (result header, set by the last kernel)
    uint32_t need_far;              // kernel 4a<false> met a record for the literal merge loop: the host re-runs kernels 4a<true>.. (exlr_abi.cu)
    uint32_t ticket_d;              // CTAs of the step's last kernel that are done: the last one stores the result header
    uint32_t pad[10];
};
static_assert(sizeof(Ctrl) == 128, "Ctrl is the 128-byte result header");

// raw indel event, kernel 1 -> kernel 4b (32 bytes, two 16-byte stores)
struct RawEv {
    uint32_t rid;       // 0xffffffff = tombstone (record did not pass the filters)
    uint32_t seq;       // index among the record's indel events, CIGAR order
    uint32_t L;         // left_consume at the op (main.rs:547-600), u32 wrapping
    uint32_t n_type;    // len | is_del << 31
    uint32_t prevL;     // L of the record's previous event (for the pair merge, main.rs:612-635)
    uint32_t pad[3];
};
static_assert(sizeof(RawEv) == 32, "RawEv");

// per-record summary written by kernel 1:  x = total_consume (main.rs:528-545), y = info
static constexpr uint32_t K1_CNT_MASK = 0x0fffffffu;   // number of indel events of the record
static constexpr uint32_t K1_MERGED = 1u << 29;        // set by kernel 4a: the record's lines come from the literal merge loop (far_list), not from its raw events
static constexpr uint32_t K1_PAIR_MERGE = 1u << 30;    // 2nd event merges with the 1st (near-edge rule, main.rs:615)
static constexpr uint32_t K1_FAR_HIT = 1u << 31;       // some adjacent pair satisfies the far-edge rule (main.rs:673-678)

// per-SA-record summary of the record's own CIGAR (kernel 3a -> 3b), the part of
// cigar_map / first_cigar_str (main.rs:214-306) the later rules read
struct SaSum {
    uint32_t S, H;          // wrapping sums of soft / hard clips (main.rs:347,353,461)
    int64_t refspan;        // D + M + '=' + X, each wrapped to u32 first (split_read_event.rs:23-28)
    int64_t ffm;            // find_first_match_pos of the rebuilt CIGAR text (utils.rs:12-42)
    uint32_t pad[2];
};
static_assert(sizeof(SaSum) == 32, "SaSum");

static constexpr uint32_t CSA_DROP = 1u << 31;         // -k cap: record skipped entirely (main.rs:311-313)
static constexpr uint32_t CSA_CNT_MASK = 0x7fffffffu;

// one alignment of a split read (SplitReadEvent, split_read_event.rs:3-11) as kernel 3b keeps it
struct Seg {
    int64_t start, end, key;    // key = find_first_match_pos(raw_cigar)
    uint32_t chrom_ref;         // tid, or sa byte offset | 1<<31 (already past "chr")
    uint32_t chrom_len;
    uint32_t clip_big;          // S > ins_clip_min || H > ins_clip_min
    uint32_t strand_neg;
};
static constexpr int kLocalSegs = 10;                  // segments kept in local memory (-k 8 fits); more -> global pool

// launch shapes, per context (host side only: the kernels never read them; they ride in the argument block so that every
// launcher sees the values of the context the batch belongs to -- nothing about a launch is process-global)
struct HostCfg {
    int sms;                // multiprocessors of the context's device
    int k1_ctas;            // persistent CTAs of kernel 1 per SM (1..4)
    int k1a_ctas;           // resident CTAs of the screen kernel per SM (fewer leave room for the SA branch beside it)
    int k1_waves;           // kernel 1 grid = SMs x CTAs/SM x waves: > 1 trades prefetch depth for dynamic balance
};

// ---- kernel argument block ----------------------------------------------------------------
struct DevBatch {
    HostCfg hc;
    // inputs
    const uint32_t* cigar; const unsigned long long* cigar_off;
    const int32_t* pos; const int32_t* tid; const uint16_t* flag; const uint8_t* mapq; const uint8_t* sa_kind;
    const uint32_t* sa_off; const uint8_t* sa_bytes;
    // reference names, "chr" stripped: name i = ref_bytes[ref_off[i] .. ref_off[i+1])
    const uint8_t* ref_bytes; const uint32_t* ref_off; int32_t n_ref;
    // scratch
    uint2* k1;              // [R] {total_consume, info}; with k1_gated only the entries of claimed records (dirty_bits) are valid
    uint32_t k1_gated;      // 1 = screened CIGAR path
    uint32_t* csa;          // [R] SA-derived line count | CSA_DROP
    uint32_t* sa_list;      // [R] ordered indices of kept records with an SA aux
    uint32_t* sa_base;      // [R] first temp slot of the SA record's events (indexed like sa_list)
    SaSum* sa_sum;          // [R] indexed like sa_list
    RawEv* raw;             // [raw_cap]: per-tile slices [0, prim_slots), then the overflow region
    uint32_t* tile_cnt;     // [R] events in each tile's slice
    uint32_t* dirty_bits;   // [R/32] one bit per record: claimed by a thread of kernel 1b (zeroed with ctrl)
    uint32_t* short_list;   // [R] k1b_claim: claimed records walked by one thread each, any order (ctrl->n_short entries)
    uint32_t* warp_list;    // [R] k1b_claim: claimed records of medium length, scanned by one warp each (ctrl->n_warp entries)
    uint2* far_list;        // [R] kernel 4a: {record, rounds of the merge loop to apply} for records the literal >2 merge loop changes
    uint32_t* long_list;    // [R] kernel 1b: claimed records too long for one thread, any order (ctrl->n_long entries)
    uint32_t* step_sum;     // [max_ops/512 + 1] kernel 1a<SUMS>: reference-consuming length of every 512-op step (long-record batches)
    uint8_t* step_flag;     // [max_ops/512 + 1] kernel 1a<SUMS>: the step holds an event candidate
    uint32_t* step_list;    // [max_ops/512 + 1] kernel 1a: the 512-op steps of the CIGAR stream that hold an event candidate, any order
    uint32_t raw_cap, prim_slots, capt_log2, slab;
    exlr_event* sa_ev;      // [max_events] SA-derived events, per record contiguous
    Seg* seg_pool; uint32_t seg_pool_cap;
    unsigned long long* scan_a; unsigned long long* scan_b;   // chained-scan tile status
    unsigned long long* scan_c;                                // ... of kernel 5a (one word per 256 events)
    Ctrl* ctrl;
    Ctrl* host_ctrl;        // the result header in mapped pinned host memory: stored by the last CTA of the step's last kernel
    // outputs
    uint32_t* line_off;     // [R+1]
    exlr_event* events;     // [max_events]
    const uint8_t* qnames; const uint32_t* qname_off;   // read names on the device (BAM batches only, else null)
    uint32_t verbose;       // 1 = kernels 5a/5b add the -v columns (utils.rs:205-223, 252-267); needs qnames
    uint32_t* text_off;     // [max_events+1] byte offset of every line (kernel 5a); null unless EXLR_OPT_DEVICE_FORMAT
    uint8_t* text;          // [text_cap] the formatted lines (kernel 5b)
    uint32_t text_cap;
    uint32_t n_reads; uint32_t max_events;
    unsigned long long* dbg;    // optional per-CTA trace (EXLR_OPT_TRACE), 4 x u64 per entry; kernel 1: {start, first data, end, tiles | scanned tiles << 32}
    uint32_t dbg_sel;           // which kernel writes the trace: 1 = kernel 1 / 1b, 2 = k0, 3 = k3a, 4 = k3b, 5 = k4a, 6 = k4b ({start, mid, end, 0})
};

struct DevParams {
    uint32_t mapq, exclude_flag, exclude_secondary, exclude_unmapped, split_only;
    uint32_t indel_min, merge_min, ins_clip_min;
    double max_pct_overlap;
    unsigned long long max_supp_alignm;
};

// ---- BAM input decoded on the device (exlr_bam.cu) -----------------------------------------
static constexpr uint32_t KI_NONE = 0xffffffffu;
struct BgzfBlock { uint32_t coff, clen, uoff, ulen, crc, pad[3]; };    // deflate data = comp[coff, coff+clen) -> U[uoff, uoff+ulen); CRC-32 of the output
static_assert(sizeof(BgzfBlock) == 32, "BgzfBlock");

struct BamCtrl {                   // device-side control block of the BAM stages (zeroed per submit)
    uint32_t bad_block;            // ~(smallest block index whose deflate stream is corrupt), 0 = none (atomicMax)
    uint32_t n_rec, corrupt;       // records found by the walk; 1 = the chain hit a corrupt record header at tail_off
    uint32_t tail_off;             // first byte of the stream the walk did not consume (a partial record, or the end)
    uint32_t n_ops, n_sa, n_qn;    // totals of the gathered batch
    uint32_t bad_rec;              // ~(smallest record index with corrupt aux data), 0 = none
    uint32_t ticket[3];
    uint32_t capped;               // 1 = more records than the batch holds (cannot happen with the bounds of exlr_bam_batch_alloc)
    uint32_t pad[4];
};
static_assert(sizeof(BamCtrl) == 64, "BamCtrl");

struct DevBam {
    uint32_t check_crc;     // 1 = kb_inflate verifies every block's CRC-32 (htslib does: a mismatch is a read error there)
    const uint8_t* comp; const BgzfBlock* blocks; uint32_t n_blocks, block_index_base;   // (index of blocks[0] in the batch's table: error reports)
    uint8_t* U; uint32_t u_begin, u_total, start_off; int32_t n_ref;    // the stream is U[u_begin, u_total); the walk starts at start_off
    uint32_t *spec, *cnt, *exitp, *kind;            // per block: speculated first record, records owned, where the chain leaves, how it stopped
    uint32_t *blk_start, *blk_base;                 // per block, verified: first record (KI_NONE: none) and index of its first record
    uint32_t* rec_start;                            // [max_reads] offset of every record's block_size word
    uint32_t *ncig, *salen, *qlen, *cig_src, *sa_src;   // per record
    uint32_t* qname_off; uint8_t* qnames;           // gathered read names (for -v and for inspection)
    unsigned long long *scan_x, *scan_y, *scan_z;   // chained-scan status words
    BamCtrl* ctrl; BamCtrl* host_ctrl;              // device block and its mirror in mapped pinned memory
    uint32_t max_reads;
};

void launch_bam_inflate(const DevBam& B, cudaStream_t st);
void launch_bam_walk(const DevBam& B, const DevBatch& D, cudaStream_t st);
uint32_t bam_scan_tiles(uint32_t max_reads);

// launchers (exlr_cigar.cu, exlr_sa.cu, exlr_order.cu)
cudaError_t configure_kernels(int device, int* sm_count_out);
size_t k1_flat_smem_bytes();
uint32_t scan_tiles(uint32_t n_reads);
uint32_t text_scan_tiles(uint32_t max_events);
void launch_k5(const DevBatch& B, cudaStream_t st);
void launch_k0(const DevBatch& B, const DevParams& P, bool walk, cudaStream_t st);
void plan_k1(DevBatch& B, int variant, uint32_t rpc, uint32_t* tiles_out);
void launch_k1(const DevBatch& B, const DevParams& P, int variant, uint32_t rpc, cudaStream_t st);
uint32_t k1a_steps(unsigned long long n_ops);
void launch_k1a(const DevBatch& B, const DevParams& P, unsigned long long n_ops, bool sums, cudaStream_t st);
void launch_k1d(const DevBatch& B, const DevParams& P, cudaStream_t st);
void launch_k1b(const DevBatch& B, const DevParams& P, unsigned long long n_ops, bool use_k1c, cudaStream_t st);
void launch_k1c(const DevBatch& B, const DevParams& P, cudaStream_t st);
void launch_k3a(const DevBatch& B, const DevParams& P, uint32_t mean_ops, cudaStream_t st);
void launch_k3b(const DevBatch& B, const DevParams& P, bool fold, cudaStream_t st);
bool k3_fold(uint32_t mean_ops);
void launch_k4a(const DevBatch& B, const DevParams& P, bool far, cudaStream_t st);
void launch_reset_tail(const DevBatch& B, cudaStream_t st);
void launch_k4b(const DevBatch& B, const DevParams& P, bool far, cudaStream_t st);

}  // namespace exlr
