// packer.hpp — lays alignment records out as the structure-of-arrays batch of include/exlr.h (exlr_batch_views).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/exlr.h"
#include "bam_reader.hpp"

namespace exlr_host {

struct PackedBatch {
    exlr_batch_views v{};
    uint64_t n = 0, ops = 0, sa = 0;
    bool keep_qnames = false;
    std::vector<char> qnames; std::vector<uint32_t> qname_off{0};

    void reset() { n = ops = sa = 0; qnames.clear(); qname_off.assign(1, 0); v.cigar_off[0] = 0; v.sa_off[0] = 0; }

    bool fits(const BamRecordView& r) const
    {
        return n < v.max_reads && ops + r.n_cigar <= v.max_ops && sa + r.sa_len <= v.max_sa_bytes;
    }
    // a record larger than an empty batch can never be packed
    bool can_ever_fit(const BamRecordView& r) const { return r.n_cigar <= v.max_ops && r.sa_len <= v.max_sa_bytes; }

    void push(const BamRecordView& r)
    {
        if (r.n_cigar) memcpy(v.cigar + ops, r.cigar, 4 * (size_t)r.n_cigar);
        ops += r.n_cigar;
        v.pos[n] = r.pos; v.tid[n] = r.tid; v.flag[n] = r.flag; v.mapq[n] = r.mapq; v.sa_kind[n] = r.sa_kind;
        if (r.sa_kind == EXLR_SA_STRING && r.sa_len) { memcpy(v.sa_bytes + sa, r.sa, r.sa_len); sa += r.sa_len; }
        n++;
        v.cigar_off[n] = ops; v.sa_off[n] = (uint32_t)sa;
        if (keep_qnames) { qnames.insert(qnames.end(), r.qname, r.qname + r.qname_len); qname_off.push_back((uint32_t)qnames.size()); }
    }
};

}  // namespace exlr_host
