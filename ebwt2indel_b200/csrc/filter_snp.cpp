// filter_snp.cpp -- bin/filter_snp: coverage filter on a .snp file.
//
// Drop-in for the reference's `filter_snp calls.snp m [M]` (/root/reference/filter_snp.cpp:17-81):
// keep the records (header line + sequence line) whose `cov:` field is >= m and, when M is given
// and non-zero, <= M; output to stdout.  The work is e2i_filter_snp in libe2i (host code), so the
// filter can also be applied to the text e2i_run returns without a round trip through a file.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>

#include "e2i.h"

int main(int argc, char **argv) {
    if (argc != 3 && argc != 4) {
        std::cout << "filter_snp calls.snp m [M]" << std::endl << std::endl
                  << "Input: a .snp file. Keep only reads with at least coverage m and at most M. Output to stdout." << std::endl;
        return 0;
    }
    std::ifstream is(argv[1], std::ios::binary);
    const std::string text((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
    char *out = nullptr;
    size_t len = 0;
    if (e2i_filter_snp(text.data(), text.size(), atoi(argv[2]), argc == 4 ? atoi(argv[3]) : 0, &out, &len) != E2I_OK) {
        std::fprintf(stderr, "filter_snp: %s\n", e2i_last_error());
        return 2;
    }
    std::fwrite(out, 1, len, stdout);
    e2i_buffer_free(out);
    return 0;
}
