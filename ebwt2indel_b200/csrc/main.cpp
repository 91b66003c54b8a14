// main.cpp -- bin/ebwt2InDel: the reference's command line over libe2i (C ABI only).
//
// Drop-in for `ebwt2InDel -1 <bwt> [-2 <bwt> | -d <da>] -o <out.snp> [-L -R -k -g -v -m -c -q -t]`
// (/root/reference/ebwt2InDel.cpp:76-103 help, :1677-1823 main).  Same flags, the same "0 means
// default" rule (:1740-1746), the same input files (raw ASCII eBWT, ASCII '0'/'1' document array),
// the same .snp bytes and the same counter lines on stdout; the progress percentages of the
// sequential loops are not reproduced.  Exit codes: help / missing file -> 0, forbidden symbol -> 1
// (dna_string.hpp:90-96), device or library failure -> 2.
#include <fcntl.h>
#include <getopt.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "e2i.h"

using std::cout;
using std::endl;
using std::string;

namespace {

[[noreturn]] void help() {
    e2i_params d;
    e2i_params_default(&d);
    cout << "ebwt2InDel [options]" << endl
         << "Options:" << endl
         << "-h          Print this help." << endl
         << "-1 <arg>    Input eBWT file (A,C,G,T,#) of first reads set (REQUIRED)." << endl
         << "-2 <arg>    Input eBWT file (A,C,G,T,#) of second reads set. If not specified, perform genotyping of first reads set." << endl
         << "            If specified, find differences (SNPs/indels) between the two reads sets." << endl
         << "-d <arg>    Input Document Array. If option -2 is not specified, this file specifies which characters from the input bwt" << endl
         << "            belong to the first (0) and which from the second (1) individual. Format: ASCII file filled with '0' and '1'." << endl
         << "-o <arg>    Output .snp file (REQUIRED)." << endl
         << "-L <arg>    Length of left-context, SNP included. Default: " << d.k_left << "." << endl
         << "-R <arg>    Length of right context, SNP excluded. Default: " << d.k_right << "." << endl
         << "-k <arg>    Minimum LCP required in clusters. Default: " << d.K << "." << endl
         << "-g <arg>    Maximum allowed gap length in indel. Default: " << d.max_gap << ". If 0, indels are disabled." << endl
         << "-v <arg>    Maximum number of non-isolated SNPs in left-contexts (excluding cntral SNP/indel). Default: " << d.max_snvs << "." << endl
         << "-m <arg>    Minimum coverage of output events. Default: " << d.mcov_out << "." << endl
         << "-c <arg>    Discard events with low-complexity right-context.  Here, low-complexity means that the context starts with a " << endl
         << "            run of <arg> equal characters. Default: length of right context (-R), minus 10." << endl
         << "-q          Maximum number of allowed variants per genomic position in each sample. If 0, there is no limit. Default: 0." << endl
         << "-t <arg>    ASCII value of terminator character. Default: " << int('#') << " (#)." << endl
         << endl
         << "\nTo run ebwt2InDel, you must first build the extended Burrows-Wheeler Transform of the input sequences." << endl
         << endl
         << "Output format: A fasta file with DNA fragments containing the variations." << endl;
    std::exit(0);
}

bool file_exists(const string &path) {
    struct stat sb;
    return ::stat(path.c_str(), &sb) == 0;
}

// Whole file into a page-locked buffer (large sequential reads; the reference reads one byte per call).
// `want` > 0 pads / truncates to that length the way the reference's DA loop does: a failed read
// leaves the previous byte in place (ebwt2InDel.cpp:1503-1508).
bool read_file(const string &path, uint8_t **buf, uint64_t *len, uint64_t want) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat sb;
    if (::fstat(fd, &sb) != 0) { ::close(fd); return false; }
    const uint64_t size = (uint64_t)sb.st_size;
    const uint64_t n = want ? want : size;
    void *p = nullptr;
    if (e2i_host_alloc(n + 16, &p) != E2I_OK) { ::close(fd); return false; }
    uint8_t *b = static_cast<uint8_t *>(p);
    uint64_t got = 0;
    const uint64_t lim = std::min(n, size);
    while (got < lim) {
        const ssize_t r = ::read(fd, b + got, (size_t)std::min<uint64_t>(lim - got, 1ull << 30));
        if (r <= 0) break;
        got += (uint64_t)r;
    }
    ::close(fd);
    if (got < n) std::memset(b + got, got ? b[got - 1] : 0, n - got);
    *buf = b;
    *len = n;
    return true;
}

void print_histogram(const e2i_stats &st) {   // ebwt2InDel.cpp:1454-1462 / 1567-1575 / 1664-1672
    uint64_t scale = 0;
    for (int i = 0; i <= 200; ++i) scale = std::max(scale, (uint64_t)st.clust_sizes[i]);
    for (int i = 0; i <= 200; ++i) {
        cout << i << (i < 10 ? "   " : (i < 100 ? "  " : " "));
        if (scale)
            for (uint64_t j = 0; j < (100 * st.clust_sizes[i]) / scale; ++j) cout << "-";
        cout << " " << st.clust_sizes[i] << endl;
    }
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 3) help();
    e2i_params p;
    std::memset(&p, 0, sizeof p);
    p.term = '#';
    string input1, input2, input_da, output;
    int opt;
    while ((opt = getopt(argc, argv, "h1:2:v:L:R:m:g:k:t:o:d:c:q:")) != -1) {
        switch (opt) {
            case 'h': help(); break;
            case '1': input1 = optarg; break;
            case 'o': output = optarg; break;
            case '2': input2 = optarg; break;
            case 'd': input_da = optarg; break;
            case 'm': p.mcov_out = atoi(optarg); break;
            case 'k': p.K = atoi(optarg); break;
            case 'g': p.max_gap = atoi(optarg); break;
            case 'L': p.k_left = atoi(optarg); break;
            case 'R': p.k_right = atoi(optarg); break;
            case 'v': p.max_snvs = atoi(optarg); break;
            case 't': p.term = (int)(char)atoi(optarg); break;
            case 'c': p.complexity = atoi(optarg); break;
            case 'q': p.max_variants_per_position = atoi(optarg); break;
            default: help(); return -1;
        }
    }
    e2i_params_resolve(&p);

    if (input1.empty() || output.empty()) help();
    if (!file_exists(input1)) {
        cout << "Error: could not find file " << input1 << endl << endl;
        help();
    }
    if (!input2.empty() && !file_exists(input2)) {
        cout << "Error: could not find file " << input2 << endl << endl;
        help();
    }
    if (!input2.empty() && !input_da.empty()) {
        cout << "Error: Document array (-d) can only be used with one input BWT file (-1)" << endl << endl;
        help();
    }

    cout << "This is ebwt2InDel." << endl;
    if (!input2.empty()) cout << "Running on two samples. Input eBWT files : " << input1 << " and " << input2 << endl;
    else if (!input_da.empty()) cout << "Running on one sample with input Document array. Input eBWT/DA files : " << input1 << " and " << input_da << endl;
    else cout << "Running on one sample (genotyping). Input eBWT file : " << input1 << endl;
    cout << "Left-extending eBWT ranges by " << p.k_left << " bases." << endl
         << "Right context length: " << p.k_right << " bases." << endl
         << "Complexity filter: " << p.complexity << endl
         << "Storing output events to file " << output << endl
         << "Minimum coverage of output events: " << p.mcov_out << endl;
    if (p.max_variants_per_position > 0) cout << "Maximum number of variants per genomic position per sample: " << p.max_variants_per_position << endl;
    else cout << "Maximum number of variants per genomic position per sample: unlimited." << endl;
    cout << endl;

    // GPUs: E2I_DEVICE=<id> (default 0) runs on one GPU; E2I_GPUS=<N> (devices 0..N-1) or E2I_DEVICES=<id,id,...>
    // runs the same job on several GPUs of this box from this one process (the reference's way to use more
    // hardware is the wrapper pebwt2InDel.sh, which splits the reads and loses cross-piece coverage).
    int device = 0;
    if (const char *dv = std::getenv("E2I_DEVICE")) device = atoi(dv);
    std::vector<int> devices;
    if (const char *dl = std::getenv("E2I_DEVICES")) {
        for (const char *q = dl; *q;) { devices.push_back(atoi(q)); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
    } else if (const char *ng = std::getenv("E2I_GPUS")) {
        for (int i = 0; i < atoi(ng); ++i) devices.push_back(i);
    }
    const bool multi = devices.size() > 1;
    uint64_t frontier_bytes = 0;
    if (const char *fb = std::getenv("E2I_FRONTIER_BYTES")) frontier_bytes = strtoull(fb, nullptr, 10);
    e2i_ctx *ctx = nullptr;
    const bool dbg = std::getenv("E2I_DEBUG") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    if (!multi) {
        if (devices.size() == 1) device = devices[0];
        if (e2i_create(device, &ctx) != E2I_OK) {
            cout << "Error: " << e2i_last_error() << endl;
            return 2;
        }
        if (frontier_bytes) e2i_set_frontier_budget(ctx, frontier_bytes);
    }

    if (dbg) std::fprintf(stderr, "[e2i] cli: context created after %.1f ms\n", since(t_start));
    const bool two = !input2.empty(), with_da = !input_da.empty();
    cout << (two ? "Phase 1/4: loading and indexing eBWTs ... " : "Phase 1/4: loading and indexing eBWT ... ") << std::flush;
    const auto t0 = std::chrono::steady_clock::now();
    uint8_t *b1 = nullptr, *b2 = nullptr, *da = nullptr;
    uint64_t n1 = 0, n2 = 0, nd = 0;
    // forbidden symbols: same message and exit code as dna_string.hpp:90-96
    auto forbidden = [&](uint8_t c) {
        cout << "Error while reading file: read forbidden character '" << (char)c << "' (ASCII code " << int((char)c) << ")." << endl
             << "Only A,C,G,T, and " << (char)p.term << " are admitted in the input BWT!" << endl
             << "If the unknown character is the terminator, you can solve the problem by adding option \"-t " << int((char)c) << "\"." << endl;
        std::exit(1);
    };
    e2i_stats st;
    std::memset(&st, 0, sizeof st);
    char *snp = nullptr;
    size_t snp_len = 0;
    int rc;
    if (!multi) {
        // one GPU: the files are streamed (reader thread -> page-locked ring -> copy stream -> counting pass)
        uint64_t bad[2] = {0, 0};
        rc = e2i_run_files(ctx, input1.c_str(), two ? input2.c_str() : nullptr, with_da ? input_da.c_str() : nullptr, &p,
                           &snp, &snp_len, &st, &n1, &n2, bad);
        if (rc == E2I_ERR_SYMBOL) {                   // fetch the byte on this error path only
            const string &f = bad[1] == 2 ? input2 : input1;
            const int fd = ::open(f.c_str(), O_RDONLY);
            uint8_t c = 0;
            if (fd >= 0 && ::pread(fd, &c, 1, (off_t)bad[0]) == 1) forbidden(c);
        }
    } else {
        if (!read_file(input1, &b1, &n1, 0) || (two && !read_file(input2, &b2, &n2, 0)) ||
            (with_da && !read_file(input_da, &da, &nd, n1))) {
            cout << "Error: could not read the input files (" << e2i_last_error() << ")" << endl;
            return 2;
        }
        rc = e2i_run_multi(devices.data(), (int)devices.size(), b1, n1, b2, n2, da, &p, frontier_bytes, &snp, &snp_len, &st);
        if (rc == E2I_ERR_SYMBOL) {
            auto check_symbols = [&](const uint8_t *b, uint64_t n) {
                for (uint64_t i = 0; i < n; ++i) {
                    const uint8_t c = b[i];
                    if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != (uint8_t)p.term) forbidden(c);
                }
            };
            check_symbols(b1, n1);
            if (two) check_symbols(b2, n2);
        }
    }
    if (rc != E2I_OK) {
        cout << "Error: " << e2i_last_error() << endl;
        return 2;
    }
    const uint64_t n = n1 + n2;
    cout << "done." << endl;
    cout << "\nPhase 2/4: " << (two ? "merging eBWTs." : "navigating suffix tree leaves.") << endl;
    if (two) cout << "Computed " << st.da_values_leaves << "/" << n << " DA values." << endl;   // :757, the leaf pass
    cout << "Computed " << st.lcp_values_leaves << "/" << n << " LCP threshold values." << endl;
    cout << "Processed " << st.leaves << " suffix-tree leaves." << endl << endl;
    cout << "Phase 3/4: computing LCP minima." << endl;
    if (two) cout << "Computed " << st.da_values << "/" << n << " DA values." << endl;
    cout << "Computed " << st.lcp_values << "/" << n << " LCP values." << endl;
    cout << "Found " << st.n_min << " LCP minima." << endl;
    cout << "Processed " << st.nodes << " suffix-tree nodes." << endl << endl;
    cout << "Phase 4/4: detecting SNPs and indels." << endl;
    cout << "Output events will be stored in " << output << endl;

    if (dbg) std::fprintf(stderr, "[e2i] cli: results ready after %.1f ms\n", since(t_start));
    FILE *f = std::fopen(output.c_str(), "wb");
    if (!f || (snp_len && std::fwrite(snp, 1, snp_len, f) != snp_len)) {
        cout << "Error: could not write " << output << endl;
        return 2;
    }
    std::fclose(f);
    if (dbg) std::fprintf(stderr, "[e2i] cli: output written after %.1f ms\n", since(t_start));

    const double avg = st.n_clusters ? double(st.clust_size) / double(st.n_clusters) : 0.0 / 0.0;
    cout << endl << "Done." << endl << "Analyzed " << st.n_clusters << " clusters." << endl
         << "Average cluster length: " << avg << "." << endl << endl;
    if (!two && !with_da) {
        cout << "Stored to file " << st.events << " events clustered in " << st.clusters_out << " clusters." << endl << endl
             << "Distribution of bases inside clusters (cluster length / number of bases inside clusters of that length): " << endl;
        print_histogram(st);
    } else {
        cout << "Distribution of bases inside clusters (cluster length / number of bases inside clusters of that length): " << endl << endl;
        print_histogram(st);
        if (with_da) cout << "\nStored to file " << 0 << " sequences clustered in " << st.clusters_out << " clusters." << endl;
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr,
                 "[e2i] n=%llu nodes=%llu leaves=%llu | device ms: h2d %.1f index %.1f leaves %.1f nodes %.1f call %.1f | wall %.3f s\n",
                 (unsigned long long)n, (unsigned long long)st.nodes, (unsigned long long)st.leaves, st.ms_h2d, st.ms_index,
                 st.ms_leaves, st.ms_nodes, st.ms_call, secs);
    e2i_buffer_free(snp);
    e2i_host_free(b1);
    e2i_host_free(b2);
    e2i_host_free(da);
    e2i_destroy(ctx);
    if (dbg) std::fprintf(stderr, "[e2i] cli: context destroyed after %.1f ms\n", since(t_start));
    return 0;
}
