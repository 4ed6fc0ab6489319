// bam_chunker.hpp — the host's share of reading a BGZF-compressed BAM when the decoding runs on the GPU (exlr_bam_*):
// read the file in large slabs straight into the batch's pinned chunk buffer and hop over the BGZF block headers
// (gzip header with the BC extra subfield: 18 bytes, ISIZE in the block's last 4 bytes; SAMv1 4.1).  Nothing is inflated
// here except the BAM header at the very start (reference names for exlr_create: reference src/main.rs:198 reads them
// through htslib's header), with zlib, once.
#pragma once
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/exlr.h"

namespace exlr_host {

class BgzfBamStream {
public:
    std::vector<std::string> ref_names;
    std::string error;
    uint64_t first_record_off = 0;     // uncompressed offset of the first record inside the first block read_blocks() returns
    uint64_t bytes_read = 0;
    int threads = 8;                   // parallel pread()s per chunk: one thread copies page cache -> pinned memory at a few GB/s only

    ~BgzfBamStream()
    {
        { std::lock_guard<std::mutex> lk(pool_mu_); pool_stop_ = true; }
        pool_cv_.notify_all();
        for (auto& t : pool_) t.join();
        if (fd_ >= 0) close(fd_);
    }

    // true: `path` is a BGZF-compressed BAM and its header has been read.  false with an empty `error`: something else
    // (SAM text, plain data): the caller uses the host reader.
    bool open(const std::string& path)
    {
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) { error = "cannot open " + path; return false; }
        // the header: inflate block after block until it is complete
        std::vector<uint8_t> hdr;
        size_t need = 12;
        for (;;) {
            top_up(64 * 1024 + 32);
            if (hdr.empty() && buf_.size() - at_ < 18) return false;
            uint32_t off, clen, ulen, total;
            const int k = parse_block(buf_.data() + at_, buf_.size() - at_, &off, &clen, &ulen, &total);
            if (k <= 0) { if (!hdr.empty()) error = "truncated BAM header"; return false; }
            std::vector<uint8_t> out(ulen ? ulen : 1);
            if (ulen) {
                z_stream zs; memset(&zs, 0, sizeof zs);
                if (inflateInit2(&zs, -15) != Z_OK) { error = "zlib"; return false; }
                zs.next_in = buf_.data() + at_ + off; zs.avail_in = clen; zs.next_out = out.data(); zs.avail_out = ulen;
                const int rc = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (rc != Z_STREAM_END || zs.avail_out != 0) { if (!hdr.empty()) error = "corrupt BGZF block in the BAM header"; return false; }
            }
            if (hdr.empty() && (ulen < 4 || memcmp(out.data(), "BAM\1", 4) != 0)) return false;      // BGZF, but not BAM (bgzipped SAM)
            const size_t before = hdr.size();
            hdr.insert(hdr.end(), out.begin(), out.begin() + ulen);
            // how much header is there?  magic, l_text, text, n_ref, then per reference l_name, name, l_ref
            for (;;) {
                if (hdr.size() < need) break;
                if (stage_ == 0) { int32_t l_text; memcpy(&l_text, hdr.data() + 4, 4); if (l_text < 0) { error = "corrupt BAM header"; return false; } pos_ = 8 + (size_t)l_text; need = pos_ + 4; stage_ = 1; }
                else if (stage_ == 1) { int32_t n; memcpy(&n, hdr.data() + pos_, 4); if (n < 0) { error = "corrupt BAM header"; return false; } n_ref_ = n; pos_ += 4; need = pos_ + (n_ref_ ? 4 : 0); stage_ = 2; }
                else if (stage_ == 2) {
                    if ((int64_t)ref_names.size() == n_ref_) { stage_ = 4; break; }
                    int32_t l_name; memcpy(&l_name, hdr.data() + pos_, 4);
                    if (l_name < 1) { error = "corrupt BAM header"; return false; }
                    name_len_ = (size_t)l_name; need = pos_ + 4 + name_len_ + 4; stage_ = 3;
                } else if (stage_ == 3) {
                    ref_names.emplace_back((const char*)hdr.data() + pos_ + 4, name_len_ - 1);
                    pos_ += 4 + name_len_ + 4; need = pos_ + 4; stage_ = 2;
                    if ((int64_t)ref_names.size() == n_ref_) { stage_ = 4; break; }
                } else break;
            }
            if (stage_ == 4) {
                // pos_ = end of the header in the uncompressed stream; this block holds it
                if (pos_ == hdr.size()) { at_ += total; first_record_off = 0; }            // records start with the next block
                else first_record_off = pos_ - before;                                     // ... inside this one: it stays in the buffer
                return true;
            }
            at_ += total;
        }
    }

    // Appends whole BGZF blocks to dst (capacity cap_bytes / cap_blocks): their bytes and their table entries (comp_off relative to
    // dst).  Returns the number of blocks; 0 at the end of the file.  A truncated last block ends the stream quietly, like a read
    // error ends the reference's loop (src/main.rs:165-168).
    size_t read_blocks(uint8_t* dst, size_t cap_bytes, exlr_bgzf_block* tab, size_t cap_blocks, size_t* bytes_out)
    {
        size_t have = 0, nb = 0, scanned = 0;
        if (carry_.size()) {                        // what the previous call read but could not return (a partial block, or blocks beyond its block capacity)
            if (carry_.size() > cap_bytes) { error = "chunk buffer too small"; *bytes_out = 0; return 0; }
            memcpy(dst, carry_.data(), carry_.size());
            have = carry_.size(); carry_.clear();
        }
        if (at_ < buf_.size()) {                    // then what open() left in its buffer
            const size_t k = std::min(buf_.size() - at_, cap_bytes - have);
            memcpy(dst + have, buf_.data() + at_, k);
            at_ += k; have += k;
            if (at_ == buf_.size()) { buf_.clear(); buf_.shrink_to_fit(); at_ = 0; }
        }
        for (;;) {
            // hop over the complete blocks that are here
            while (nb < cap_blocks) {
                uint32_t off, clen, ulen, total;
                const int k = parse_block(dst + scanned, have - scanned, &off, &clen, &ulen, &total);
                if (k < 0) { error = "not a BGZF block"; eof_ = true; bad_ = true; have = scanned; break; }
                if (k == 0) break;
                uint32_t crc; memcpy(&crc, dst + scanned + total - 8, 4);
                tab[nb++] = exlr_bgzf_block{(uint32_t)(scanned + off), clen, ulen, crc};
                scanned += total;
            }
            if (eof_ || nb == cap_blocks || have == cap_bytes || at_ < buf_.size()) break;
            // read on, straight into the caller's (pinned) buffer: the rest of its capacity, cut into slices read side by side
            const size_t n = read_parallel(dst + have, cap_bytes - have);
            if (n == 0) { eof_ = true; break; }
            have += n;
        }
        if (have > scanned && !bad_) carry_.assign(dst + scanned, dst + have);     // starts the next chunk (at the end of the file: a truncated block, never completed)
        *bytes_out = scanned;
        return nb;
    }


private:
    int fd_ = -1;
    std::vector<uint8_t> buf_, carry_;
    size_t at_ = 0;
    uint64_t fpos_ = 0;
    bool eof_ = false, bad_ = false;
    int stage_ = 0; size_t pos_ = 0, name_len_ = 0; int64_t n_ref_ = 0;

    // at least `want` bytes behind at_ in the header buffer, unless the file ends first
    void top_up(size_t want)
    {
        if (at_ > 0) { buf_.erase(buf_.begin(), buf_.begin() + (ptrdiff_t)at_); at_ = 0; }
        while (buf_.size() < want && !eof_) {
            const size_t base = buf_.size();
            buf_.resize(base + (1u << 20));
            const ssize_t n = ::pread(fd_, buf_.data() + base, 1u << 20, (off_t)fpos_);
            buf_.resize(base + (n > 0 ? (size_t)n : 0));
            if (n <= 0) eof_ = true; else { bytes_read += (size_t)n; fpos_ += (size_t)n; }
        }
    }

    // ---- the readers: `threads - 1` workers that live as long as the stream (a chunk is read every few milliseconds: starting
    // 31 threads per chunk cost a quarter of the read itself) plus the calling thread.  A job is a range of the file cut into
    // slices; whoever is free takes the next slice.
    std::vector<std::thread> pool_;
    std::mutex pool_mu_;
    std::condition_variable pool_cv_, done_cv_;
    bool pool_stop_ = false;
    int pool_busy_ = 0;                   // workers inside read_slices() (guarded by pool_mu_): a new job is only set up at 0
    uint64_t job_id_ = 0;                 // bumped per job (guarded by pool_mu_)
    uint8_t* job_dst_ = nullptr; size_t job_want_ = 0, job_slices_ = 0; uint64_t job_fpos_ = 0;
    std::atomic<size_t> job_next_{0}, job_left_{0};
    std::vector<size_t> job_got_;

    void read_slices()
    {
        for (;;) {
            const size_t t = job_next_.fetch_add(1);
            if (t >= job_slices_) return;
            const size_t a = job_want_ * t / job_slices_, b = job_want_ * (t + 1) / job_slices_;
            size_t done = 0;
            while (a + done < b) {
                const ssize_t n = ::pread(fd_, job_dst_ + a + done, b - a - done, (off_t)(job_fpos_ + a + done));
                if (n <= 0) break;
                done += (size_t)n;
            }
            job_got_[t] = done;
            if (job_left_.fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(pool_mu_); done_cv_.notify_all(); }
        }
    }
    void worker()
    {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(pool_mu_);
                pool_cv_.wait(lk, [&] { return pool_stop_ || job_id_ != seen; });
                if (pool_stop_) return;
                seen = job_id_;
                pool_busy_++;
            }
            read_slices();
            { std::lock_guard<std::mutex> lk(pool_mu_); pool_busy_--; }
            done_cv_.notify_all();
        }
    }

    // up to `want` bytes from the file position on; short only at the end of the file
    size_t read_parallel(uint8_t* dst, size_t want)
    {
        const size_t kSlice = 2u << 20;
        const size_t slices = std::max<size_t>(1, (want + kSlice - 1) / kSlice);
        const int nt = (int)std::min<size_t>((size_t)std::max(threads, 1), slices);
        while ((int)pool_.size() < nt - 1) pool_.emplace_back([this] { worker(); });
        {
            std::unique_lock<std::mutex> lk(pool_mu_);
            done_cv_.wait(lk, [&] { return pool_busy_ == 0; });     // (a worker of the previous job may still be on its way out)
            job_dst_ = dst; job_want_ = want; job_slices_ = slices; job_fpos_ = fpos_;
            job_got_.assign(slices, 0);
            job_next_.store(0); job_left_.store(slices);
            job_id_++;
        }
        pool_cv_.notify_all();
        read_slices();
        { std::unique_lock<std::mutex> lk(pool_mu_); done_cv_.wait(lk, [&] { return job_left_.load() == 0 && pool_busy_ == 0; }); }
        size_t total = 0;
        for (size_t t = 0; t < slices; t++) {                // contiguous prefix (a short slice means the file ended inside it)
            const size_t a = want * t / slices, b = want * (t + 1) / slices;
            total += job_got_[t];
            if (job_got_[t] < b - a) break;
        }
        fpos_ += total; bytes_read += total;
        return total;
    }

    // 1: a complete block at p (offset / length of its DEFLATE data, ISIZE, total size); 0: not all there yet; -1: not BGZF
    static int parse_block(const uint8_t* p, size_t n, uint32_t* off, uint32_t* clen, uint32_t* ulen, uint32_t* total)
    {
        if (n < 18) return 0;
        if (p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) return -1;
        const uint32_t xlen = (uint32_t)p[10] | ((uint32_t)p[11] << 8);
        if (n < 12 + (size_t)xlen) return 0;
        int bsize = -1;
        for (uint32_t i = 12; i + 4 <= 12 + xlen;) {
            const uint32_t slen = (uint32_t)p[i + 2] | ((uint32_t)p[i + 3] << 8);
            if (p[i] == 'B' && p[i + 1] == 'C' && slen == 2 && i + 6 <= 12 + xlen) bsize = (int)((uint32_t)p[i + 4] | ((uint32_t)p[i + 5] << 8));
            i += 4 + slen;
        }
        if (bsize < 0) return -1;
        const uint32_t t = (uint32_t)bsize + 1;
        if (t < xlen + 20) return -1;
        if (n < t) return 0;
        *off = 12 + xlen; *clen = t - xlen - 20; *total = t;
        memcpy(ulen, p + t - 4, 4);
        return *ulen <= 65536u ? 1 : -1;
    }
};

}  // namespace exlr_host
