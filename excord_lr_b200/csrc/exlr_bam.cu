// exlr_bam.cu — BAM input decoded on the device: what htslib does for the reference between the file and the record loop
// (bam::Reader::from_path / set_threads / bam.read(&mut record), reference src/main.rs:137,155,158), as CUDA kernels:
//
//   kb_inflate   one warp per BGZF block: RFC 1951 DEFLATE (stored / fixed / dynamic Huffman + LZ77) straight into the
//                chunk's uncompressed stream; the host only hops over the 18-byte block headers
//   kb_spec      one warp per block: the first plausible record header in the block, then the record chain from there
//                (block_size hops) -- a speculation, because where a block enters the record stream is only known serially
//   kb_verify    one warp: walks the blocks in order and accepts every speculation that starts exactly where the true
//                chain arrives (else re-walks that block itself): exact for any input, parallel for every sane one
//   kb_emit      record start offsets, block by block
//   kb_fields    one thread per record: fixed fields, aux walk for SA (first occurrence, like bam_aux_get) and CG
//                (htslib's long-CIGAR restore, bam_tag2cigar inside bam_read1), qname
//   kb_offsets   chained scan: CIGAR-op / SA-byte / qname-byte offsets of every record
//   kb_copy      one warp per record: CIGAR ops, SA string and qname into the structure-of-arrays batch the event kernels read
//
// Byte / integer work bound by HBM and by the serial nature of Huffman decoding inside one block; parallelism is across the
// thousands of BGZF blocks of a chunk.  No tensor cores.
#include "exlr_common.cuh"

namespace exlr {

// ======================================================================================
// kb_inflate
// ======================================================================================
static constexpr int KI_WARPS = 8;             // BGZF blocks in flight per CTA
static constexpr int KI_LL_BITS = 10;          // literal/length codes up to this many bits resolve with one table look-up
static constexpr int KI_D_BITS = 9;
static constexpr uint32_t KI_RING_LINES = 4;   // 128-byte lines of compressed input held per warp

struct __align__(16) InflWarp {
    uint16_t ll[1 << KI_LL_BITS];              // (symbol << 4) | code length, indexed by the next bits of the stream; 0 = longer code
    uint16_t dd[1 << KI_D_BITS];
    uint16_t ll_sym[288], d_sym[32];           // symbols ordered by (code length, symbol): canonical decode of the longer codes
    uint16_t ll_cnt[16], d_cnt[16];            // codes per length
    uint16_t cl[128];                          // the code-length code (at most 7 bits)
    uint16_t cl_sym[20], cl_cnt[16];
    uint32_t nc[16], so[16], cw[16];           // table building scratch: next code / sorted offset / count per length
    uint8_t lens[320];
    uint32_t ring[KI_RING_LINES * 32];         // the compressed stream, prefetched (BitReader)
};

// All 32 lanes of the warp run the decoder redundantly (same data, same control flow: one instruction stream); what differs per
// lane is the byte it holds of the pending output.  The compressed stream is prefetched into a small per-warp ring in shared
// memory with cp.async (four 128-byte lines: the line being read and three ahead), so no register ever waits for global memory:
// a line is requested ~384 bytes of input before it is read.
struct BitReader {
    const uint32_t* line0;                     // 128-byte aligned start of the stream's first line
    uint32_t* ring;                            // [KI_RING_LINES * 32] words in shared memory
    uint32_t line, widx, bc;                   // current line, next word in it, valid bits in bb
    unsigned long long bb;
    __device__ __forceinline__ void fetch_line(uint32_t l)     // every lane copies one word of line l into its ring slot
    {
        const uint32_t lane = threadIdx.x & 31;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (l % KI_RING_LINES) * 32 + lane);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n\tcp.async.commit_group;" ::"r"(dst), "l"(line0 + (size_t)l * 32 + lane) : "memory");
    }
    __device__ __forceinline__ void init(const uint8_t* p)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        line0 = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)127);
        line = 0; widx = (uint32_t)(a & 127) >> 2; bb = 0; bc = 0;
        // (a re-init behind a stored block: lines requested for the old position may still be in flight to the same ring slots, and
        // two cp.async writes to one address have no order of their own -- drain them before the ring is filled again)
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (uint32_t l = 0; l < KI_RING_LINES; l++) fetch_line(l);
        asm volatile("cp.async.wait_group %0;" ::"n"(KI_RING_LINES - 1) : "memory");
        __syncwarp();
        refill(); refill();
        const uint32_t skip = (uint32_t)(a & 3) * 8u;
        bb >>= skip; bc -= skip;
        refill();
    }
    __device__ __forceinline__ void refill()    // one word if there is room for it; afterwards bc >= 33 (callers consume < 32 bits between calls)
    {
        if (bc <= 32u) {
            const uint32_t w = ring[(line % KI_RING_LINES) * 32 + widx];
            bb |= (unsigned long long)w << bc; bc += 32u;
            if (++widx == 32u) {
                // the line just finished frees its slot: request the line KI_RING_LINES ahead there, then make sure the next one landed
                __syncwarp();
                fetch_line(line + KI_RING_LINES);
                line++; widx = 0;
                asm volatile("cp.async.wait_group %0;" ::"n"(KI_RING_LINES - 1) : "memory");
                __syncwarp();
            }
        }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(uint32_t n) { bb >>= n; bc -= n; }
    __device__ __forceinline__ uint32_t get(uint32_t n) { const uint32_t v = peek(n); drop(n); return v; }
    // address of the next unread byte once the reader is byte aligned
    __device__ __forceinline__ const uint8_t* byte_ptr() const
    {
        return reinterpret_cast<const uint8_t*>(line0) + ((size_t)line * 32 + widx) * 4 - bc / 8;
    }
};

// Canonical Huffman tables from code lengths (RFC 1951 3.2.2), built by the whole warp.  Returns false for an over-subscribed
// set of lengths.  lut: one look-up for codes of at most `bits` bits; sym/cnt: canonical decode for the longer ones.
__device__ __forceinline__ bool build_table(InflWarp& S, const uint8_t* lens, uint32_t n, uint16_t* lut, uint32_t bits, uint16_t* sym, uint16_t* cnt)
{
    const uint32_t lane = threadIdx.x & 31;
    if (lane < 16) S.cw[lane] = 0;
    for (uint32_t i = lane; i < (1u << bits); i += 32) lut[i] = 0;
    __syncwarp();
    for (uint32_t s = lane; s < n; s += 32) { const uint32_t l = lens[s]; if (l) atomicAdd(&S.cw[l], 1u); }
    __syncwarp();
    int left = 1; uint32_t code = 0, at = 0; bool ok = true;
    for (uint32_t l = 1; l <= 15; l++) {                                     // every lane, same values
        const uint32_t c = S.cw[l];
        left = (left << 1) - (int)c;
        if (left < 0) ok = false;
        if (lane == 0) { S.nc[l] = code; S.so[l] = at; cnt[l] = (uint16_t)c; }
        code = (code + c) << 1; at += c;
    }
    if (lane == 0) cnt[0] = 0;
    __syncwarp();
    if (!ok) return false;
    for (uint32_t base = 0; base < n; base += 32) {                          // codes go to the symbols in symbol order
        const uint32_t s = base + lane, l = s < n ? lens[s] : 0u;
        const uint32_t same = __match_any_sync(0xffffffffu, l), rank = __popc(same & ((1u << lane) - 1u));
        if (l) {
            const uint32_t c = S.nc[l] + rank;
            sym[S.so[l] + rank] = (uint16_t)s;
            if (l <= bits) {
                const uint32_t rev = __brev(c) >> (32u - l);                 // the stream carries Huffman codes most significant bit first
                const uint16_t ent = (uint16_t)((s << 4) | l | (s < 256u ? 0x2000u : 0u));   // bit 13: a literal (only ever tested in the literal/length table)
                for (uint32_t k = rev; k < (1u << bits); k += 1u << l) lut[k] = ent;
            }
        }
        __syncwarp();
        if (l && lane == (uint32_t)(__ffs((int)same) - 1)) { const uint32_t k = __popc(same); S.nc[l] += k; S.so[l] += k; }
        __syncwarp();
    }
    return true;
}

// one symbol: table look-up, or bit by bit for codes longer than the table (puff-style canonical decode); false = invalid code
__device__ __forceinline__ bool decode_sym(BitReader& br, const uint16_t* lut, uint32_t bits, const uint16_t* sym, const uint16_t* cnt, uint32_t* out)
{
    const uint32_t e = lut[br.peek(bits)];
    if (e) { br.drop(e & 15u); *out = (e >> 4) & 0x1ffu; return true; }
    uint32_t code = 0, first = 0, index = 0;
    for (uint32_t l = 1; l <= 15; l++) {
        code |= (uint32_t)(br.bb >> (l - 1)) & 1u;
        const uint32_t c = cnt[l];
        if (code - first < c) { *out = sym[index + (code - first)]; br.drop(l); return true; }
        index += c; first = (first + c) << 1; code <<= 1;
    }
    return false;
}

// order in which the code length code lengths are stored (RFC 1951 3.2.7)
__constant__ uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// ---- CRC-32 of a block's output (gzip / BGZF footer; htslib's bgzf_read_block verifies it and the reference's reader then
// returns an error).  Every lane takes the CRC of one slice of the output with the byte table; the slices are combined the
// way zlib's crc32_combine does: crc(A || B) = crc(A) * x^(8 |B|) + crc(B) over GF(2)[x] mod the CRC polynomial.
static constexpr uint32_t KI_CRC_POLY = 0xedb88320u;
__device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b)      // a * b mod P, bit 31 = x^0 (zlib multmodp)
{
    uint32_t p = 0;
    for (uint32_t m = 0x80000000u; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ KI_CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8 n) mod P from the table of x^(2^k) (zlib x2nmodp with k = 3)
__device__ __forceinline__ uint32_t gf2_x8n(const uint32_t* x2n, uint32_t n)
{
    uint32_t p = 0x80000000u;
    for (uint32_t k = 3; n; n >>= 1, k++) if (n & 1u) p = gf2_mulmod(x2n[k & 31u], p);
    return p;
}

__global__ void __launch_bounds__(KI_WARPS * 32) kb_inflate(const uint8_t* __restrict__ comp, const BgzfBlock* __restrict__ blocks, uint32_t n_blocks,
                                                             uint32_t index_base, uint8_t* U, BamCtrl* ctrl, uint32_t check_crc)
{
    __shared__ InflWarp s_all[KI_WARPS];
    __shared__ uint32_t s_crc_tab[256], s_x2n[32];
    if (check_crc) {                                                     // (block-uniform)
        uint32_t c = threadIdx.x;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ KI_CRC_POLY : c >> 1;
        if (threadIdx.x < 256) s_crc_tab[threadIdx.x] = c;
        if (threadIdx.x == 0) { uint32_t p = 0x40000000u; s_x2n[0] = p; for (int n = 1; n < 32; n++) s_x2n[n] = p = gf2_mulmod(p, p); }
        __syncthreads();
    }
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    InflWarp& S = s_all[w];
    const uint32_t b = blockIdx.x * KI_WARPS + w;
    if (b >= n_blocks) return;
    const BgzfBlock blk = blocks[b];
    if (blk.ulen == 0) return;                                               // (the EOF marker block)
    uint8_t* out = U + blk.uoff;
    const uint32_t ulen = blk.ulen;
    BitReader br; br.ring = S.ring; br.init(comp + blk.coff);
    uint32_t o = 0, ps = 0, pend = 0;                                        // output position; pending literals are [ps, o), one per lane
    const uint32_t oa = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 31u);  // (oa + o) & 31 = the lane that holds output byte o
    const uint32_t my_o = (lane - oa) & 31u, wend = (32u - oa) & 31u;       // ... i.e. o & 31 == my_o; a 32-byte window ends where o & 31 == wend
    bool ok = true;
    // pending literal bytes sit in the lane (address & 31) of their 32-byte window and go out as one 32-byte store
    auto flush = [&]() {
        const uint32_t w0 = (oa + ps) & ~31u, mine = w0 + lane;             // window and my byte, counted from out - oa
        if (mine >= oa + ps && mine < oa + o && o <= ulen) (out - oa)[mine] = (uint8_t)pend;
        ps = o;
    };
    for (bool last = false; ok && !last;) {
        br.refill();
        last = br.get(1) != 0;
        const uint32_t type = br.get(2);
        if (type == 0) {                                                     // stored
            br.drop(br.bc & 7u);
            br.refill();
            const uint32_t len = br.get(16);
            br.refill();
            const uint32_t nlen = br.get(16);
            if ((len ^ 0xffffu) != nlen || o + len > ulen) { ok = false; break; }
            flush();
            const uint8_t* src = br.byte_ptr();
            for (uint32_t i = lane; i < len; i += 32) out[o + i] = __ldg(src + i);
            o += len; ps = o;
            br.init(src + len);
            continue;
        }
        if (type == 3) { ok = false; break; }
        uint32_t hlit = 288, hdist = 30;
        if (type == 1) {                                                     // fixed code (RFC 1951 3.2.6)
            for (uint32_t s = lane; s < 288; s += 32) S.lens[s] = s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8));
            if (lane < 30) S.lens[288 + lane] = 5;
            __syncwarp();
        } else {                                                             // dynamic code (3.2.7)
            hlit = br.get(5) + 257; hdist = br.get(5) + 1;
            const uint32_t hclen = br.get(4) + 4;
            if (lane < 19) S.lens[lane] = 0;
            __syncwarp();
            for (uint32_t i = 0; i < hclen; i++) {
                br.refill();
                const uint32_t v = br.get(3);
                if (lane == 0) S.lens[kClOrder[i]] = (uint8_t)v;
            }
            __syncwarp();
            if (!build_table(S, S.lens, 19, S.cl, 7, S.cl_sym, S.cl_cnt)) { ok = false; break; }
            __syncwarp();
            const uint32_t total = hlit + hdist;
            uint32_t i = 0, prev = 0;
            while (i < total) {
                br.refill();
                uint32_t sym;
                if (!decode_sym(br, S.cl, 7, S.cl_sym, S.cl_cnt, &sym)) { ok = false; break; }
                uint32_t rep = 1, val = sym;
                if (sym == 16) { if (i == 0) { ok = false; break; } val = prev; rep = 3 + br.get(2); }
                else if (sym == 17) { val = 0; rep = 3 + br.get(3); }
                else if (sym == 18) { val = 0; rep = 11 + br.get(7); }
                if (i + rep > total) { ok = false; break; }
                for (uint32_t k = lane; k < rep; k += 32) S.lens[i + k] = (uint8_t)val;
                i += rep; prev = val;
            }
            if (!ok) break;
            __syncwarp();
            if (S.lens[256] == 0) { ok = false; break; }                     // no end-of-block code
            // the distance lengths follow the literal/length lengths: move them to a fixed place
            uint32_t dl = lane < hdist ? S.lens[hlit + lane] : 0u;
            __syncwarp();
            if (lane < 32) S.lens[288 + lane] = (uint8_t)dl;
            __syncwarp();
        }
        if (!build_table(S, S.lens, hlit, S.ll, KI_LL_BITS, S.ll_sym, S.ll_cnt)) { ok = false; break; }
        if (!build_table(S, S.lens + 288, hdist, S.dd, KI_D_BITS, S.d_sym, S.d_cnt)) { ok = false; break; }
        __syncwarp();
        for (;;) {                                                           // the symbols of this deflate block
            br.refill();
            {   // literals, the common case, straight from the table: one look-up, the byte parks in its lane
                const uint32_t e = S.ll[br.peek(KI_LL_BITS)];
                if (e & 0x2000u) {
                    br.drop(e & 15u);
                    if ((o & 31u) == my_o) pend = e >> 4;
                    o++;
                    if ((o & 31u) == wend) { if (o > ulen) { ok = false; break; } flush(); }
                    continue;
                }
            }
            uint32_t sym;
            if (!decode_sym(br, S.ll, KI_LL_BITS, S.ll_sym, S.ll_cnt, &sym)) { ok = false; break; }
            if (sym < 256u) {                                                // literal with a code longer than the table
                if (lane == ((oa + o) & 31u)) pend = sym;
                o++;
                if (((oa + o) & 31u) == 0u) { if (o > ulen) { ok = false; break; } flush(); }
                continue;
            }
            if (sym == 256u) break;                                          // end of block
            const uint32_t c = sym - 257u;
            if (c >= 29u) { ok = false; break; }
            uint32_t len;
            if (c < 8u) len = c + 3u;
            else if (c == 28u) len = 258u;
            else { const uint32_t eb = (c >> 2) - 1u; len = ((4u + (c & 3u)) << eb) + 3u + br.get(eb); }
            br.refill();
            uint32_t dc;
            if (!decode_sym(br, S.dd, KI_D_BITS, S.d_sym, S.d_cnt, &dc) || dc >= 30u) { ok = false; break; }
            uint32_t dist;
            if (dc < 4u) dist = dc + 1u;
            else { const uint32_t eb = (dc >> 1) - 1u; dist = ((2u + (dc & 1u)) << eb) + 1u + br.get(eb); }
            if (dist > o || o + len > ulen) { ok = false; break; }
            flush();
            __syncwarp();
            // byte o+i = byte o+i-dist; for i >= dist that is a byte of this very copy: o-dist + (i mod dist) names the same
            // value among the bytes that already exist, so all lanes copy at once
            const uint8_t* src = out + o - dist;
            if (dist >= len) { for (uint32_t i = lane; i < len; i += 32) out[o + i] = src[i]; }
            else { for (uint32_t i = lane; i < len; i += 32) out[o + i] = src[i % dist]; }
            __syncwarp();
            o += len; ps = o;
        }
    }
    flush();                                                                 // (stores nothing once o is past ulen)
    if (ok && o == ulen && check_crc) {
        __syncwarp();
        const uint32_t L = (ulen + 31u) / 32u, a0 = min(lane * L, ulen), a1 = min(a0 + L, ulen);
        uint32_t c = 0xffffffffu;
        for (uint32_t i = a0; i < a1; i++) c = s_crc_tab[(c ^ out[i]) & 0xffu] ^ (c >> 8);
        c = a1 > a0 ? ~c : 0u;                                               // crc32 of my slice (of nothing: 0)
        const uint32_t g = gf2_x8n(s_x2n, a1 - a0);                          // x^(8 |slice|)
        uint32_t acc = 0;
        for (int i = 0; i < 32; i++) acc = gf2_mulmod(__shfl_sync(0xffffffffu, g, i), acc) ^ __shfl_sync(0xffffffffu, c, i);
        if (acc != blk.crc) ok = false;
    }
    if (!ok || o != ulen) { if (lane == 0) atomicMax(&ctrl->bad_block, ~(index_base + b)); }
}

// ======================================================================================
// record walk
// ======================================================================================
__device__ __forceinline__ uint32_t ld32u(const uint8_t* U, uint32_t p)      // unaligned little-endian u32 (U is padded)
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(U) + (p >> 2);
    return __funnelshift_r(w[0], w[1], (p & 3u) * 8u);
}
__device__ __forceinline__ uint32_t ld16u(const uint8_t* U, uint32_t p) { return (uint32_t)U[p] | ((uint32_t)U[p + 1] << 8); }

// Could a BAM record start at p?  (SAMv1 4.2: block_size, refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, l_seq, next_refID,
// next_pos, tlen, read_name ...)  Only a filter for the speculation: kb_verify decides.
__device__ __forceinline__ bool plausible_record(const uint8_t* U, uint32_t p, uint32_t total, int32_t n_ref)
{
    if ((unsigned long long)p + 36ull > total) return false;
    const uint32_t bs = ld32u(U, p);
    if (bs < 32u || bs > (1u << 29)) return false;
    const int32_t tid = (int32_t)ld32u(U, p + 4), pos = (int32_t)ld32u(U, p + 8), ntid = (int32_t)ld32u(U, p + 24), npos = (int32_t)ld32u(U, p + 28);
    if (tid < -1 || tid >= n_ref || ntid < -1 || ntid >= n_ref || pos < -1 || npos < -1) return false;
    const uint32_t l_name = U[p + 12], n_cig = ld16u(U, p + 16);
    const int32_t l_seq = (int32_t)ld32u(U, p + 20);
    if (l_name == 0 || l_seq < 0) return false;
    const unsigned long long need = 32ull + l_name + 4ull * n_cig + ((unsigned long long)l_seq + 1) / 2 + (unsigned long long)l_seq;
    if (need > bs) return false;
    if ((unsigned long long)p + 36ull + l_name <= total && U[p + 36 + l_name - 1] != 0) return false;    // read_name is NUL terminated
    return true;
}

// the chain from p while it stays below `be`: number of records, where it leaves, and why it stopped
// (0 = left the block, 1 = the record at *x is not complete inside the stream, 2 = corrupt block_size at *x)
__device__ __forceinline__ uint32_t walk_chain(const uint8_t* U, uint32_t p, uint32_t be, uint32_t total, uint32_t* x, uint32_t* kind)
{
    uint32_t n = 0; *kind = 0;
    while (p < be) {
        if ((unsigned long long)p + 4ull > total) { *kind = 1; break; }
        const uint32_t bs = ld32u(U, p);
        if (bs < 32u || bs > 0x7fffffffu) { *kind = 2; break; }
        if ((unsigned long long)p + 4ull + bs > total) { *kind = 1; break; }
        n++; p += 4u + bs;
    }
    *x = p;
    return n;
}

__global__ void __launch_bounds__(256) kb_spec(DevBam B)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B.n_blocks) return;
    const BgzfBlock blk = B.blocks[b];
    const uint32_t bs = blk.uoff, be = blk.uoff + blk.ulen;
    uint32_t spec = KI_NONE;
    if (be > B.start_off && blk.ulen) {
        if (bs <= B.start_off) spec = B.start_off;                           // the stream position the walk starts from is a fact
        else {
            for (uint32_t p0 = bs; p0 < be && spec == KI_NONE; p0 += 32) {
                const uint32_t p = p0 + lane;
                const uint32_t m = __ballot_sync(0xffffffffu, p < be && plausible_record(B.U, p, B.u_total, B.n_ref));
                if (m) spec = p0 + (uint32_t)(__ffs((int)m) - 1);
            }
        }
    }
    uint32_t n = 0, x = be, kind = 0;
    if (spec != KI_NONE) n = walk_chain(B.U, spec, be, B.u_total, &x, &kind);
    if (lane == 0) { B.spec[b] = spec; B.cnt[b] = n; B.exitp[b] = x; B.kind[b] = kind; }
}

__global__ void __launch_bounds__(32) kb_verify(DevBam B)
{
    const uint32_t lane = threadIdx.x;
    uint32_t cur = B.start_off, running = 0, stop = 0;                       // stop: 0 running, 1 tail, 2 corrupt
    for (uint32_t b0 = 0; b0 < B.n_blocks; b0 += 32) {
        const uint32_t b = b0 + lane;
        uint32_t v_be = 0, v_spec = KI_NONE, v_cnt = 0, v_exit = 0, v_kind = 0;
        if (b < B.n_blocks) { const BgzfBlock k = B.blocks[b]; v_be = k.uoff + k.ulen; v_spec = B.spec[b]; v_cnt = B.cnt[b]; v_exit = B.exitp[b]; v_kind = B.kind[b]; }
        uint32_t my_start = KI_NONE, my_base = 0;
        const uint32_t nb = min(32u, B.n_blocks - b0);
        for (uint32_t i = 0; i < nb; i++) {
            const uint32_t be = __shfl_sync(0xffffffffu, v_be, i);
            uint32_t start = KI_NONE, n = 0;
            if (!stop && cur < be) {
                uint32_t nxt, kind;
                if (__shfl_sync(0xffffffffu, v_spec, i) == cur) { n = __shfl_sync(0xffffffffu, v_cnt, i); nxt = __shfl_sync(0xffffffffu, v_exit, i); kind = __shfl_sync(0xffffffffu, v_kind, i); }
                else n = walk_chain(B.U, cur, be, B.u_total, &nxt, &kind);  // the speculation missed: this block is walked here
                start = n ? cur : KI_NONE;
                cur = nxt; stop = kind;
            }
            if (lane == i) { my_start = start; my_base = running; }
            running += n;
        }
        if (b < B.n_blocks) { B.blk_start[b] = my_start; B.blk_base[b] = my_base; }
    }
    if (lane == 0) {
        BamCtrl* c = B.ctrl;
        c->n_rec = running; c->tail_off = stop ? cur : min(cur, B.u_total); c->corrupt = stop == 2u;
        if (running > B.max_reads) c->capped = 1;
    }
}

__global__ void __launch_bounds__(128) kb_emit(DevBam B)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B.n_blocks) return;
    uint32_t p = B.blk_start[b];
    if (p == KI_NONE) return;
    const BgzfBlock blk = B.blocks[b];
    const uint32_t be = blk.uoff + blk.ulen;
    uint32_t at = B.blk_base[b];
    while (p < be && at < B.max_reads) {
        if ((unsigned long long)p + 4ull > B.u_total) break;
        const uint32_t bs = ld32u(B.U, p);
        if (bs < 32u || bs > 0x7fffffffu || (unsigned long long)p + 4ull + bs > B.u_total) break;
        B.rec_start[at++] = p;
        p += 4u + bs;
    }
}

// ======================================================================================
// gather: records -> structure of arrays
// ======================================================================================
__device__ __forceinline__ uint32_t aux_size(uint32_t t)
{
    switch (t) { case 'A': case 'c': case 'C': return 1; case 's': case 'S': return 2; case 'i': case 'I': case 'f': return 4; default: return 0; }
}

__global__ void __launch_bounds__(256) kb_fields(DevBam B, DevBatch D)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = min(B.ctrl->n_rec, B.max_reads);
    if (r >= n) return;
    const uint8_t* U = B.U;
    const uint32_t p = B.rec_start[r] + 4u, bs = ld32u(U, p - 4u);
    const int32_t tid = (int32_t)ld32u(U, p), pos = (int32_t)ld32u(U, p + 4);
    const uint32_t bmn = ld32u(U, p + 8), fnc = ld32u(U, p + 12), l_seq = ld32u(U, p + 16);
    const uint32_t l_name = bmn & 0xffu, n_cigar = fnc & 0xffffu;
    const_cast<int32_t*>(D.tid)[r] = tid; const_cast<int32_t*>(D.pos)[r] = pos;
    const_cast<uint16_t*>(D.flag)[r] = (uint16_t)(fnc >> 16); const_cast<uint8_t*>(D.mapq)[r] = (uint8_t)((bmn >> 8) & 0xffu);
    const unsigned long long seq_bytes = ((unsigned long long)l_seq + 1ull) / 2ull + (unsigned long long)l_seq;
    const unsigned long long fixed = 32ull + l_name + 4ull * n_cigar + seq_bytes;
    uint32_t sa_kind = 0, sa_src = 0, sa_len = 0, ncig = n_cigar, cig_src = p + 32u + l_name, qlen = l_name ? l_name - 1u : 0u;
    bool bad = fixed > bs || l_name == 0;
    if (!bad) {
        // aux walk: SA (first occurrence, like bam_aux_get) and CG (first occurrence)
        uint32_t o = (uint32_t)fixed, cg_at = 0, cg_n = 0; bool cg_seen = false;
        while (o + 3u <= bs) {
            const uint32_t t0 = U[p + o], t1 = U[p + o + 1], ty = U[p + o + 2];
            unsigned long long v = (unsigned long long)o + 3ull, len;
            if (ty == 'Z' || ty == 'H') {
                unsigned long long e = v;
                while (e < bs && U[p + e]) e++;
                if (e >= bs) { bad = true; break; }
                len = e - v + 1ull;
            } else if (ty == 'B') {
                if (v + 5ull > bs) { bad = true; break; }
                const uint32_t sub = U[p + v], es = aux_size(sub), cnt = ld32u(U, p + (uint32_t)v + 1u);
                if (!es) { bad = true; break; }
                len = 5ull + (unsigned long long)es * cnt;
                if (t0 == 'C' && t1 == 'G' && !cg_seen) { cg_seen = true; if (sub == 'I' || sub == 'i') { cg_at = p + (uint32_t)v + 5u; cg_n = cnt; } }
            } else { len = aux_size(ty); if (!len) { bad = true; break; } }
            if (v + len > bs) { bad = true; break; }
            if (t0 == 'S' && t1 == 'A' && sa_kind == 0) {
                if (ty == 'Z') { sa_kind = EXLR_SA_STRING; sa_src = p + (uint32_t)v; sa_len = (uint32_t)len - 1u; }
                else sa_kind = EXLR_SA_OTHER;
            }
            o = (uint32_t)(v + len);
        }
        // long-CIGAR convention (SAMv1 4.2.2), the test of htslib's bam_tag2cigar: first op <l_seq>S, the record placed, and a
        // CG:B,I array of at least n_cigar and fewer than 2^29 ops
        if (!bad && cg_at && n_cigar > 0 && tid >= 0 && pos >= 0 && cg_n >= n_cigar && cg_n < (1u << 29)) {
            const uint32_t c0 = ld32u(U, cig_src);
            if ((c0 & 15u) == 4u && (c0 >> 4) == l_seq) { cig_src = cg_at; ncig = cg_n; }
        }
    }
    if (bad) { atomicMax(&B.ctrl->bad_rec, ~r); sa_kind = 0; sa_len = 0; ncig = 0; qlen = 0; }
    const_cast<uint8_t*>(D.sa_kind)[r] = (uint8_t)sa_kind;
    B.ncig[r] = ncig; B.salen[r] = sa_len; B.qlen[r] = qlen; B.cig_src[r] = cig_src; B.sa_src[r] = sa_src;
}

// offsets of every record's CIGAR ops / SA bytes / qname bytes: three chained scans in one pass
__global__ void __launch_bounds__(SCAN_THREADS) kb_offsets(DevBam B, DevBatch D)
{
    __shared__ uint32_t s_tile, s_warp[16];
    if (threadIdx.x == 0) s_tile = atomicAdd(&B.ctrl->ticket[0], 1u);
    __syncthreads();
    const uint32_t tile = s_tile, n = min(B.ctrl->n_rec, B.max_reads);
    const uint32_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tile >= n_tiles) {
        if (n == 0 && tile == 0 && threadIdx.x == 0) {
            const_cast<unsigned long long*>(D.cigar_off)[0] = 0; const_cast<uint32_t*>(D.sa_off)[0] = 0; B.qname_off[0] = 0;
            B.ctrl->n_ops = 0; B.ctrl->n_sa = 0; B.ctrl->n_qn = 0;
        }
        return;
    }
    const uint32_t r0 = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t a[SCAN_ITEMS], s[SCAN_ITEMS], q[SCAN_ITEMS], ta = 0, ts = 0, tq = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint32_t r = r0 + i;
        a[i] = r < n ? B.ncig[r] : 0u; s[i] = r < n ? B.salen[r] : 0u; q[i] = r < n ? B.qlen[r] : 0u;
        ta += a[i]; ts += s[i]; tq += q[i];
    }
    uint32_t ga, gs, gq;
    uint32_t xa = tile_excl_scan(B.scan_x, tile, ta, s_warp, &ga);
    __syncthreads();
    uint32_t xs = tile_excl_scan(B.scan_y, tile, ts, s_warp, &gs);
    __syncthreads();
    uint32_t xq = tile_excl_scan(B.scan_z, tile, tq, s_warp, &gq);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint32_t r = r0 + i;
        if (r < n) { const_cast<unsigned long long*>(D.cigar_off)[r] = xa; const_cast<uint32_t*>(D.sa_off)[r] = xs; B.qname_off[r] = xq; }
        xa += a[i]; xs += s[i]; xq += q[i];
    }
    if (threadIdx.x == 0 && tile == n_tiles - 1) {
        const_cast<unsigned long long*>(D.cigar_off)[n] = ga; const_cast<uint32_t*>(D.sa_off)[n] = gs; B.qname_off[n] = gq;
        B.ctrl->n_ops = ga; B.ctrl->n_sa = gs; B.ctrl->n_qn = gq;
    }
}

// one warp per record: CIGAR ops (4-byte units at any alignment), SA string, read name
__global__ void __launch_bounds__(256) kb_copy(DevBam B, DevBatch D)
{
    const uint32_t lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    const uint32_t n = min(B.ctrl->n_rec, B.max_reads);
    const uint8_t* U = B.U;
    uint32_t* cig = const_cast<uint32_t*>(D.cigar);
    uint8_t* sab = const_cast<uint8_t*>(D.sa_bytes);
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nw) {
        const uint32_t nc = B.ncig[r], src = B.cig_src[r];
        const unsigned long long co = D.cigar_off[r];
        for (uint32_t i = lane; i < nc; i += 32) cig[co + i] = ld32u(U, src + 4u * i);
        const uint32_t sl = B.salen[r], ss = B.sa_src[r], so = D.sa_off[r];
        for (uint32_t i = lane; i < sl; i += 32) sab[so + i] = U[ss + i];
        const uint32_t ql = B.qlen[r], qs = B.rec_start[r] + 36u, qo = B.qname_off[r];
        for (uint32_t i = lane; i < ql; i += 32) B.qnames[qo + i] = U[qs + i];
    }
}

// the decode header to mapped pinned host memory (what exlr_bam_extract waits for)
__global__ void __launch_bounds__(32) kb_header(DevBam B)
{
    if (threadIdx.x < sizeof(BamCtrl) / 16) reinterpret_cast<uint4*>(B.host_ctrl)[threadIdx.x] = reinterpret_cast<const uint4*>(B.ctrl)[threadIdx.x];
    __threadfence_system();
}

// ======================================================================================
// launchers
// ======================================================================================
void launch_bam_inflate(const DevBam& B, cudaStream_t st)
{
    if (!B.n_blocks) return;
    kb_inflate<<<(B.n_blocks + KI_WARPS - 1) / KI_WARPS, KI_WARPS * 32, 0, st>>>(B.comp, B.blocks, B.n_blocks, B.block_index_base, B.U, B.ctrl, B.check_crc);
}

void launch_bam_walk(const DevBam& B, const DevBatch& D, cudaStream_t st)
{
    const uint32_t nb = B.n_blocks ? B.n_blocks : 1u;
    kb_spec<<<(nb * 32 + 255) / 256, 256, 0, st>>>(B);
    kb_verify<<<1, 32, 0, st>>>(B);
    kb_emit<<<(nb + 127) / 128, 128, 0, st>>>(B);
    // the record count lives on the device: the grids cover what the chunk can hold at most, capped at a few resident waves
    const uint32_t cap = (uint32_t)D.hc.sms * 8u;
    const uint32_t bound = min(B.max_reads, (B.u_total - B.u_begin) / 36u + 1u);
    kb_fields<<<(bound + 255) / 256, 256, 0, st>>>(B, D);
    kb_offsets<<<(bound + SCAN_TILE - 1) / SCAN_TILE, SCAN_THREADS, 0, st>>>(B, D);
    kb_copy<<<min((bound + 7u) / 8u, cap * 4u), 256, 0, st>>>(B, D);
    kb_header<<<1, 32, 0, st>>>(B);
}

uint32_t bam_scan_tiles(uint32_t max_reads) { return (max_reads + SCAN_TILE - 1) / SCAN_TILE + 1; }

}  // namespace exlr
