// bam_reader.hpp — minimal multi-threaded BGZF/BAM reader for the record packer.
//
// Plays the role rust-htslib/htslib play for the reference (`bam::Reader::from_path`, `set_threads`, `bam.read(&mut record)`:
// reference src/main.rs:137,155,158): BGZF blocks are inflated on worker threads (zlib), records are walked touching only
// what the hot path reads — the fixed header (refID, pos, mapq, flag), the CIGAR (with the CG:B,I long-CIGAR restore htslib
// performs inside bam_read1), the SA aux field and, for -v, the read name.  SEQ/QUAL are skipped.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace exlr_host {

struct BamRecordView {
    int32_t tid, pos;
    uint16_t flag;
    uint8_t mapq;
    const uint32_t* cigar; uint32_t n_cigar;     // points into the decompressed stream (or into cg_restore)
    uint8_t sa_kind;                             // 0 none, 1 Z string, 2 other aux type (include/exlr.h EXLR_SA_*)
    const uint8_t* sa; uint32_t sa_len;
    const char* qname; uint32_t qname_len;       // without the trailing NUL
};

class BamReader {
public:
    std::vector<std::string> ref_names;
    std::vector<int64_t> ref_lens;
    std::string error;

    ~BamReader() { if (fp_ && fp_ != stdin) fclose(fp_); }

    // BAM, or SAM text (plain or BGZF-compressed): htslib auto-detects the same way and the reference's Makefile feeds it
    // a .sam (reference Makefile:73-74).
    bool open(const std::string& path, int threads)
    {
        fp_ = path == "-" ? stdin : fopen(path.c_str(), "rb");     // "-": a stream (no seeking anywhere in this reader)
        if (!fp_) { error = "cannot open " + path; return false; }
        threads_ = threads < 1 ? 1 : threads;
        uint8_t magic[2] = {0, 0};
        const size_t got = fread(magic, 1, 2, fp_);
        pre_.assign(magic, magic + got);                           // pushed back: the input may be a pipe
        plain_ = !(got == 2 && magic[0] == 31 && magic[1] == 139);
        if (!fill()) { if (error.empty()) error = "empty or truncated input"; return false; }
        if (!plain_ && avail() >= 4 && memcmp(cur(), "BAM\1", 4) == 0) return read_header();
        sam_ = true;
        return read_sam_header();
    }

    // Next record; false at EOF or on a read error (the reference stops silently on a read error too: src/main.rs:165-168).
    bool next(BamRecordView& r)
    {
        if (sam_) return next_sam(r);
        for (;;) {
            if (avail() >= 4) {
                uint32_t bs; memcpy(&bs, cur(), 4);
                if (bs < 32) { error = "corrupt BAM record"; return false; }
                if (avail() >= 4 + (size_t)bs) return parse(cur() + 4, bs, r);
            }
            if (!fill()) { if (avail() != 0 && error.empty()) error = "truncated BAM"; return false; }
        }
    }

private:
    FILE* fp_ = nullptr;
    int threads_ = 1;
    std::vector<uint8_t> pre_;        // bytes read while sniffing the format, served again by rd()

    size_t rd(void* dst, size_t n)
    {
        size_t k = 0;
        if (!pre_.empty()) {
            k = std::min(n, pre_.size());
            memcpy(dst, pre_.data(), k);
            pre_.erase(pre_.begin(), pre_.begin() + (ptrdiff_t)k);
            if (k == n) return k;
        }
        return k + fread((uint8_t*)dst + k, 1, n - k, fp_);
    }
    std::vector<uint8_t> buf_;        // decompressed stream window
    size_t off_ = 0;                  // read cursor inside buf_
    bool eof_ = false, plain_ = false, sam_ = false;
    std::vector<uint32_t> cg_restore_;
    std::vector<uint8_t> comp_;       // compressed superblock
    struct Blk { size_t coff, clen, uoff, ulen; };

    const uint8_t* cur() const { return buf_.data() + off_; }
    size_t avail() const { return buf_.size() - off_; }

    // Read up to ~256 BGZF blocks per thread, inflate them on threads_ workers, append to the stream window.
    // A truncated or corrupt block ends the stream after the blocks before it (the reference, too, keeps what it had read).
    bool fill()
    {
        if (eof_) return false;
        if (off_ > 0) { buf_.erase(buf_.begin(), buf_.begin() + (ptrdiff_t)off_); off_ = 0; }
        if (plain_) {
            const size_t base = buf_.size(), chunk = 4u << 20;
            buf_.resize(base + chunk);
            const size_t n = rd(buf_.data() + base, chunk);
            buf_.resize(base + n);
            if (n == 0) { eof_ = true; return false; }
            return true;
        }
        comp_.clear();
        std::vector<Blk> blks;
        size_t utotal = 0;
        const size_t kMaxBlocks = 256 * (size_t)threads_;
        while (blks.size() < kMaxBlocks) {
            uint8_t h[18];
            size_t n = rd(h, 18);
            if (n == 0) { eof_ = true; break; }
            if (n < 18 || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { error = "not a BGZF block"; eof_ = true; break; }
            // extra subfields: find BC
            uint16_t xlen; memcpy(&xlen, h + 10, 2);
            if (xlen < 6) { error = "BGZF block without BC field"; eof_ = true; break; }
            std::vector<uint8_t> extra(xlen);
            memcpy(extra.data(), h + 12, xlen < 6 ? xlen : 6);
            if (xlen > 6 && rd(extra.data() + 6, xlen - 6) != (size_t)xlen - 6) { error = "truncated BGZF header"; eof_ = true; break; }
            int bsize = -1;
            for (size_t i = 0; i + 4 <= extra.size();) {
                uint16_t slen; memcpy(&slen, &extra[i + 2], 2);
                if (extra[i] == 'B' && extra[i + 1] == 'C' && slen == 2 && i + 6 <= extra.size()) { uint16_t b; memcpy(&b, &extra[i + 4], 2); bsize = b; }
                i += 4 + slen;
            }
            if (bsize < 0) { error = "BGZF block without BC field"; eof_ = true; break; }
            const size_t total = (size_t)bsize + 1, head = 12 + (size_t)xlen;
            if (total < head + 8) { error = "corrupt BGZF block size"; eof_ = true; break; }
            const size_t body = total - head;                     // deflate data + crc32 + isize
            const size_t at = comp_.size();
            comp_.resize(at + body);
            if (rd(comp_.data() + at, body) != body) { error = "truncated BGZF block"; eof_ = true; break; }
            uint32_t isize; memcpy(&isize, comp_.data() + at + body - 4, 4);
            if (isize > 65536) { error = "corrupt BGZF isize"; eof_ = true; break; }
            blks.push_back({at, body - 8, utotal, isize});
            utotal += isize;
        }
        if (blks.empty()) return false;
        const size_t base = buf_.size();
        buf_.resize(base + utotal);
        std::atomic<size_t> next{0};
        std::atomic<bool> ok{true};
        auto work = [&]() {
            z_stream zs;
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= blks.size()) break;
                const Blk& b = blks[i];
                if (b.ulen == 0) continue;
                memset(&zs, 0, sizeof zs);
                if (inflateInit2(&zs, -15) != Z_OK) { ok = false; break; }
                zs.next_in = comp_.data() + b.coff; zs.avail_in = (uInt)b.clen;
                zs.next_out = buf_.data() + base + b.uoff; zs.avail_out = (uInt)b.ulen;
                const int rc = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (rc != Z_STREAM_END || zs.avail_out != 0) { ok = false; break; }
            }
        };
        const int nt = (int)std::min<size_t>((size_t)threads_, blks.size());
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        if (!ok) { error = "BGZF inflate failed"; eof_ = true; buf_.resize(base); return false; }
        return true;
    }

    bool need(size_t n) { while (avail() < n) if (!fill()) return false; return true; }

    bool read_header()
    {
        if (!need(12)) { error = "truncated BAM header"; return false; }
        if (memcmp(cur(), "BAM\1", 4) != 0) { error = "not a BAM file (SAM text and CRAM are not supported by this reader)"; return false; }
        int32_t l_text; memcpy(&l_text, cur() + 4, 4);
        if (l_text < 0 || !need(12 + (size_t)l_text)) { error = "truncated BAM header"; return false; }
        int32_t n_ref; memcpy(&n_ref, cur() + 8 + l_text, 4);
        off_ += 12 + (size_t)l_text;
        for (int i = 0; i < n_ref; i++) {
            if (!need(4)) { error = "truncated BAM header"; return false; }
            int32_t l_name; memcpy(&l_name, cur(), 4);
            if (l_name < 1 || !need(8 + (size_t)l_name)) { error = "truncated BAM header"; return false; }
            ref_names.emplace_back((const char*)cur() + 4, (size_t)l_name - 1);
            int32_t l_ref; memcpy(&l_ref, cur() + 4 + l_name, 4);
            ref_lens.push_back(l_ref);
            off_ += 8 + (size_t)l_name;
        }
        return true;
    }

    // one text line [cur(), cur()+len) without the newline; false at EOF.  A last line without '\n' is still returned.
    bool sam_line(size_t* len)
    {
        for (;;) {
            const uint8_t* p = cur();
            const void* nl = avail() ? memchr(p, '\n', avail()) : nullptr;
            if (nl) { *len = (size_t)((const uint8_t*)nl - p); return true; }
            if (!fill()) { if (avail() == 0) return false; *len = avail(); return true; }
        }
    }
    void sam_consume(size_t len) { off_ += len; if (avail() && *cur() == '\n') off_++; }

    bool read_sam_header()
    {
        size_t len;
        while (avail() || !eof_) {
            if (!avail() && !fill()) break;
            if (*cur() != '@') break;
            if (!sam_line(&len)) break;
            std::string line((const char*)cur(), len);
            if (!line.empty() && line.back() == '\r') line.pop_back();
            if (line.rfind("@SQ", 0) == 0) {
                std::string name; int64_t ln = 0;
                size_t b = 3;
                while (b < line.size()) {
                    size_t e = line.find('\t', b + 1); if (e == std::string::npos) e = line.size();
                    const std::string f = line.substr(b + 1, e - b - 1);
                    if (f.rfind("SN:", 0) == 0) name = f.substr(3);
                    else if (f.rfind("LN:", 0) == 0) ln = atoll(f.c_str() + 3);
                    b = e;
                }
                ref_names.push_back(name); ref_lens.push_back(ln);
            }
            sam_consume(len);
        }
        return true;
    }

    bool next_sam(BamRecordView& r)
    {
        size_t len;
        for (;;) {
            if (!sam_line(&len)) return false;
            if (len == 0 || *cur() == '@') { sam_consume(len); continue; }
            break;
        }
        const char* p = (const char*)cur();
        size_t n = len;
        if (n && p[n - 1] == '\r') n--;
        const char* f[12]; size_t fl[12]; int nf = 0; size_t b = 0, aux_at = n;
        for (size_t i = 0; i <= n && nf < 11; i++) {
            if (i == n || p[i] == '\t') { f[nf] = p + b; fl[nf] = i - b; nf++; b = i + 1; if (nf == 11) aux_at = b <= n ? b : n; }
        }
        if (nf < 11) { error = "corrupt SAM record"; return false; }
        auto num = [](const char* s, size_t l, long long* out) { if (!l) return false; char* e; std::string t(s, l); *out = strtoll(t.c_str(), &e, 10); return *e == 0; };
        long long flag, pos, mapq;
        if (!num(f[1], fl[1], &flag) || !num(f[3], fl[3], &pos) || !num(f[4], fl[4], &mapq)) { error = "corrupt SAM record"; return false; }
        r.qname = f[0]; r.qname_len = (uint32_t)fl[0];
        r.flag = (uint16_t)flag; r.pos = (int32_t)(pos - 1); r.mapq = (uint8_t)mapq;
        r.tid = -1;
        if (!(fl[2] == 1 && f[2][0] == '*')) {
            for (size_t i = 0; i < ref_names.size(); i++) if (ref_names[i].size() == fl[2] && memcmp(ref_names[i].data(), f[2], fl[2]) == 0) { r.tid = (int32_t)i; break; }
            if (r.tid < 0) { error = "SAM record on a contig that is not in the header"; return false; }
        }
        cg_restore_.clear();
        if (!(fl[5] == 1 && f[5][0] == '*')) {
            uint64_t v = 0; bool have = false;
            for (size_t i = 0; i < fl[5]; i++) {
                const char c = f[5][i];
                if (c >= '0' && c <= '9') { v = v * 10 + (uint64_t)(c - '0'); have = true; if (v >= (1ull << 28)) { error = "CIGAR length out of range"; return false; } continue; }
                static const char ops[] = "MIDNSHP=XB";
                const char* q = (const char*)memchr(ops, c, 10);
                if (!q || !have) { error = "corrupt CIGAR"; return false; }
                cg_restore_.push_back((uint32_t)(v << 4) | (uint32_t)(q - ops));
                v = 0; have = false;
            }
            if (have) { error = "corrupt CIGAR"; return false; }
        }
        r.cigar = cg_restore_.data(); r.n_cigar = (uint32_t)cg_restore_.size();
        r.sa_kind = 0; r.sa = nullptr; r.sa_len = 0;
        for (size_t i = aux_at; i < n;) {
            size_t e = i; while (e < n && p[e] != '\t') e++;
            if (e - i >= 5 && p[i] == 'S' && p[i + 1] == 'A' && p[i + 2] == ':' && p[i + 4] == ':' && r.sa_kind == 0) {
                if (p[i + 3] == 'Z') { r.sa_kind = 1; r.sa = (const uint8_t*)p + i + 5; r.sa_len = (uint32_t)(e - i - 5); }
                else r.sa_kind = 2;
            }
            i = e + 1;
        }
        sam_consume(len);
        return true;
    }

    static size_t aux_size(uint8_t type)
    {
        switch (type) { case 'A': case 'c': case 'C': return 1; case 's': case 'S': return 2; case 'i': case 'I': case 'f': return 4; default: return 0; }
    }

    bool parse(const uint8_t* p, uint32_t bs, BamRecordView& r)
    {
        int32_t tid, pos; uint32_t bin_mq_nl, flag_nc, l_seq;
        memcpy(&tid, p, 4); memcpy(&pos, p + 4, 4); memcpy(&bin_mq_nl, p + 8, 4); memcpy(&flag_nc, p + 12, 4); memcpy(&l_seq, p + 16, 4);
        const uint32_t l_name = bin_mq_nl & 0xff, n_cigar = flag_nc & 0xffff;
        r.tid = tid; r.pos = pos; r.mapq = (uint8_t)((bin_mq_nl >> 8) & 0xff); r.flag = (uint16_t)(flag_nc >> 16);
        size_t o = 32;
        const size_t seq_bytes = ((size_t)l_seq + 1) / 2 + (size_t)l_seq;       // size_t: a corrupt l_seq must not wrap into range
        const size_t fixed = o + l_name + 4 * (size_t)n_cigar + seq_bytes;
        if (fixed > bs || l_name == 0) { error = "corrupt BAM record"; return false; }
        r.qname = (const char*)p + o; r.qname_len = l_name - 1; o += l_name;
        r.cigar = (const uint32_t*)(p + o); r.n_cigar = n_cigar; o += 4 * (size_t)n_cigar;   // 4-byte alignment is not guaranteed: copied below
        o += seq_bytes;
        // aux walk: SA (first occurrence, like bam_aux_get) and CG
        r.sa_kind = 0; r.sa = nullptr; r.sa_len = 0;
        const uint8_t* cg = nullptr; uint32_t cg_n = 0; bool cg_seen = false;
        while (o + 3 <= bs) {
            const uint8_t t0 = p[o], t1 = p[o + 1], ty = p[o + 2];
            size_t v = o + 3, len;
            if (ty == 'Z' || ty == 'H') { size_t e = v; while (e < bs && p[e]) e++; if (e >= bs) { error = "corrupt aux"; return false; } len = e - v + 1; }
            else if (ty == 'B') {
                if (v + 5 > bs) { error = "corrupt aux"; return false; }
                const size_t es = aux_size(p[v]); uint32_t cnt; memcpy(&cnt, p + v + 1, 4);
                if (!es) { error = "corrupt aux"; return false; }
                len = 5 + es * (size_t)cnt;
                if (t0 == 'C' && t1 == 'G' && !cg_seen) { cg_seen = true; if (p[v] == 'I' || p[v] == 'i') { cg = p + v + 5; cg_n = cnt; } }
            } else { len = aux_size(ty); if (!len) { error = "corrupt aux"; return false; } }
            if (v + len > bs) { error = "corrupt aux"; return false; }
            if (t0 == 'S' && t1 == 'A' && r.sa_kind == 0) {
                if (ty == 'Z') { r.sa_kind = 1; r.sa = p + v; r.sa_len = (uint32_t)(len - 1); }
                else r.sa_kind = 2;
            }
            o = v + len;
        }
        // CIGAR bytes are only 4-byte aligned by luck: always hand out an aligned copy
        const uint8_t* src = (const uint8_t*)r.cigar; uint32_t n = n_cigar;
        // long-CIGAR convention (SAM spec 4.2.2).  Same test as htslib's bam_tag2cigar, which runs inside bam_read1 before the
        // reference ever sees the record: first op is <l_seq>S, the record is placed, a CG:B,I (or B,i) aux exists and its
        // count is at least n_cigar and below 2^29 (otherwise htslib keeps the CIGAR as it stands).
        if (cg && n_cigar > 0 && tid >= 0 && pos >= 0 && cg_n >= n_cigar && cg_n < (1u << 29)) {
            uint32_t c0; memcpy(&c0, src, 4);
            if ((c0 & 15) == 4 && (c0 >> 4) == l_seq) { src = cg; n = cg_n; }
        }
        cg_restore_.resize(n);
        if (n) memcpy(cg_restore_.data(), src, 4 * (size_t)n);
        r.cigar = cg_restore_.data(); r.n_cigar = n;
        off_ += 4 + (size_t)bs;
        return true;
    }
};

}  // namespace exlr_host
