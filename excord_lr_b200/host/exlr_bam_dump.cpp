// exlr_bam_dump — reader/packer check tool (no GPU): BAM -> the packed structure-of-arrays batch, written as a flat binary
// file that tests compare with the batch the BAM was generated from.
//   exlr_bam_dump in.bam out.bin [threads]
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "bam_chunker.hpp"
#include "bam_reader.hpp"

using namespace exlr_host;

template <class T> static void put(FILE* f, const std::vector<T>& v) { uint64_t n = v.size(); fwrite(&n, 8, 1, f); if (n) fwrite(v.data(), sizeof(T), n, f); }

int main(int argc, char** argv)
{
    if (argc >= 3 && std::string(argv[1]) == "--chunker") {
        // the host side of the GPU BAM decoder (bam_chunker.hpp): prints what it sees, chunk by chunk
        //   exlr_bam_dump --chunker in.bam [cap_bytes] [cap_blocks] [threads]
        BgzfBamStream bs;
        const size_t cap = argc > 3 ? (size_t)atoll(argv[3]) : (8u << 20), capb = argc > 4 ? (size_t)atoll(argv[4]) : 4096;
        bs.threads = argc > 5 ? atoi(argv[5]) : 4;
        if (!bs.open(argv[2])) { printf("not-bgzf-bam %s\n", bs.error.c_str()); return 1; }
        printf("refs %zu first_record_off %llu\n", bs.ref_names.size(), (unsigned long long)bs.first_record_off);
        std::vector<uint8_t> buf(cap); std::vector<exlr_bgzf_block> tab(capb);
        unsigned long long nb_total = 0, u_total = 0, c_total = 0, crc = 0;
        for (;;) {
            size_t bytes = 0;
            const size_t nb = bs.read_blocks(buf.data(), cap, tab.data(), capb, &bytes);
            if (!nb) break;
            for (size_t i = 0; i < nb; i++) { u_total += tab[i].ulen; c_total += tab[i].comp_len; for (uint32_t k = 0; k < tab[i].comp_len; k += 97) crc = crc * 131 + buf[tab[i].comp_off + k]; }
            nb_total += nb;
        }
        printf("blocks %llu ulen %llu clen %llu sample %llu\n", nb_total, u_total, c_total, crc);
        return 0;
    }
    if (argc < 3) { fputs("usage: exlr_bam_dump in.bam out.bin [threads]\n", stderr); return 2; }
    BamReader rd;
    if (!rd.open(argv[1], argc > 3 ? atoi(argv[3]) : 4)) { fprintf(stderr, "%s\n", rd.error.c_str()); return 1; }
    std::vector<uint32_t> cigar, sa_off{0}; std::vector<uint64_t> cigar_off{0}; std::vector<int32_t> pos, tid;
    std::vector<uint16_t> flag; std::vector<uint8_t> mapq, sa_kind, sa_bytes; std::vector<char> qn; std::vector<uint32_t> qoff{0};
    BamRecordView r;
    while (rd.next(r)) {
        cigar.insert(cigar.end(), r.cigar, r.cigar + r.n_cigar); cigar_off.push_back(cigar.size());
        pos.push_back(r.pos); tid.push_back(r.tid); flag.push_back(r.flag); mapq.push_back(r.mapq); sa_kind.push_back(r.sa_kind);
        if (r.sa_kind == 1) sa_bytes.insert(sa_bytes.end(), r.sa, r.sa + r.sa_len);
        sa_off.push_back((uint32_t)sa_bytes.size());
        qn.insert(qn.end(), r.qname, r.qname + r.qname_len); qoff.push_back((uint32_t)qn.size());
    }
    if (!rd.error.empty()) fprintf(stderr, "stopped: %s\n", rd.error.c_str());
    FILE* f = fopen(argv[2], "wb");
    if (!f) return 1;
    std::vector<char> names; std::vector<uint32_t> noff{0};
    for (auto& s : rd.ref_names) { names.insert(names.end(), s.begin(), s.end()); noff.push_back((uint32_t)names.size()); }
    put(f, names); put(f, noff); put(f, cigar); put(f, cigar_off); put(f, pos); put(f, tid); put(f, flag); put(f, mapq); put(f, sa_kind);
    put(f, sa_off); put(f, sa_bytes); put(f, qn); put(f, qoff);
    fclose(f);
    return rd.error.empty() ? 0 : 4;
}
