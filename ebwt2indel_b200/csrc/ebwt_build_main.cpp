// ebwt_build_main.cpp -- bin/ebwt_build: reads (FASTA / FASTQ / one read per line) -> eBWT (+ document array),
// the input format of ebwt2InDel.  Stands in for the external construction step the reference's README
// prescribes (BCR_LCP_GSA / egap, /root/reference/README.md:38, 91-92) so that the tool chain is self-contained:
//
//   ebwt_build -i reads.fa [-r] -o reads.ebwt                         then  ebwt2InDel -1 reads.ebwt -o out.snp
//   ebwt_build -i a.fa -j b.fa [-r] -o merged.ebwt -d merged.da       then  ebwt2InDel -1 merged.ebwt -d merged.da -o out.snp
//
// -r appends the reverse complement of every read (after the forward reads of its set).  All reads must have
// the same length.  Calls libe2i's e2i_ebwt_build (GPU, BCR-style column insertion); no CPU fallback.
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "e2i.h"

namespace {

[[noreturn]] void usage() {
    std::cout << "ebwt_build [options]\n"
                 "-i <arg>    reads of the first set: FASTA, FASTQ or one read per line (REQUIRED)\n"
                 "-j <arg>    reads of a second set (merged eBWT + document array, for ebwt2InDel -d)\n"
                 "-o <arg>    output eBWT file (REQUIRED)\n"
                 "-d <arg>    output document array (ASCII '0'/'1'); required with -j\n"
                 "-r          add the reverse complement of every read\n"
                 "-t <arg>    ASCII value of the terminator character. Default: 35 (#)\n";
    std::exit(0);
}

// appends the reads of `path` to `rows` (upper-cased); returns false on a format error
bool load_reads(const std::string &path, std::vector<std::string> &rows) {
    std::ifstream in(path);
    if (!in) { std::cout << "Error: could not read " << path << std::endl; return false; }
    std::string line, seq;
    int mode = -1;   // 0 FASTA, 1 FASTQ, 2 plain
    long fq = 0;
    auto flush = [&] { if (!seq.empty()) { rows.push_back(seq); seq.clear(); } };
    while (std::getline(in, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
        if (line.empty()) continue;
        if (mode < 0) mode = line[0] == '>' ? 0 : line[0] == '@' ? 1 : 2;
        if (mode == 0) {
            if (line[0] == '>') flush(); else seq += line;
        } else if (mode == 1) {
            if (fq % 4 == 1) rows.push_back(line);
            ++fq;
        } else {
            rows.push_back(line);
        }
    }
    flush();
    return true;
}

std::string revcomp(const std::string &s) {
    std::string r(s.rbegin(), s.rend());
    for (char &c : r) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
    return r;
}

}  // namespace

int main(int argc, char **argv) {
    std::string in1, in2, out, out_da;
    bool rc_flag = false;
    int term = '#', opt;
    while ((opt = getopt(argc, argv, "hi:j:o:d:rt:")) != -1) {
        switch (opt) {
            case 'i': in1 = optarg; break;
            case 'j': in2 = optarg; break;
            case 'o': out = optarg; break;
            case 'd': out_da = optarg; break;
            case 'r': rc_flag = true; break;
            case 't': term = atoi(optarg); break;
            default: usage();
        }
    }
    if (in1.empty() || out.empty() || (!in2.empty() && out_da.empty())) usage();
    std::vector<std::string> rows;
    auto add_set = [&](const std::string &path) {
        const size_t first = rows.size();
        if (!load_reads(path, rows)) return false;
        for (size_t i = first; i < rows.size(); ++i)
            for (char &c : rows[i]) c = (char)std::toupper((unsigned char)c);
        if (rc_flag) { const size_t last = rows.size(); for (size_t i = first; i < last; ++i) rows.push_back(revcomp(rows[i])); }
        return true;
    };
    if (!add_set(in1)) return 2;
    const uint64_t second_from = rows.size();
    if (!in2.empty() && !add_set(in2)) return 2;
    if (rows.empty()) { std::cout << "Error: no reads in " << in1 << std::endl; return 2; }
    const size_t L = rows[0].size();
    for (size_t i = 0; i < rows.size(); ++i)
        if (rows[i].size() != L) { std::cout << "Error: read " << i << " has length " << rows[i].size() << ", expected " << L << " (all reads must have the same length)" << std::endl; return 2; }
    const uint64_t m = rows.size();
    std::vector<uint8_t> mat(m * L);
    for (uint64_t i = 0; i < m; ++i) std::memcpy(mat.data() + i * L, rows[i].data(), L);
    rows.clear();
    rows.shrink_to_fit();
    std::vector<uint8_t> bwt(m * (L + 1)), da(out_da.empty() ? 0 : m * (L + 1));
    int device = 0;
    if (const char *dv = std::getenv("E2I_DEVICE")) device = atoi(dv);
    e2i_ctx *ctx = nullptr;
    if (e2i_create(device, &ctx) != E2I_OK) { std::cout << "Error: " << e2i_last_error() << std::endl; return 2; }
    const int rc = e2i_ebwt_build(ctx, mat.data(), m, (uint32_t)L, second_from, (uint8_t)term, bwt.data(), da.empty() ? nullptr : da.data());
    if (rc != E2I_OK) { std::cout << "Error: " << e2i_last_error() << std::endl; e2i_destroy(ctx); return rc == E2I_ERR_SYMBOL ? 1 : 2; }
    e2i_destroy(ctx);
    auto dump = [](const std::string &path, const std::vector<uint8_t> &v) {
        FILE *f = std::fopen(path.c_str(), "wb");
        const bool ok = f && std::fwrite(v.data(), 1, v.size(), f) == v.size();
        if (f) std::fclose(f);
        if (!ok) std::cout << "Error: could not write " << path << std::endl;
        return ok;
    };
    if (!dump(out, bwt) || (!out_da.empty() && !dump(out_da, da))) return 2;
    std::cout << "eBWT of " << m << " reads of length " << L << " (" << bwt.size() << " symbols) written to " << out << std::endl;
    return 0;
}
