// host_selftest -- CPU-only access to the pure pieces of the host pipeline (no GPU call is made):
//   host_selftest count            < bytes   -> number of '\n'
//   host_selftest cut FMT WINDOW   < bytes   -> the guessed record boundary near the end (FMT 1 = FASTA, 2 = FASTQ)
//   host_selftest canon FMT        < record  -> the record's canonical bytes (needletail write_fasta / write_fastq)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "demux_host.hpp"

int main(int argc, char** argv) {
    std::vector<uint8_t> in;
    uint8_t buf[1 << 16];
    for (size_t n; (n = fread(buf, 1, sizeof buf, stdin)) > 0;) in.insert(in.end(), buf, buf + n);
    if (argc >= 2 && !strcmp(argv[1], "count")) {
        // also at odd offsets: the vector loop and its scalar tail must agree wherever the buffer starts
        const unsigned long long c = sabreur::detail::host_count_newlines(in.data(), in.size());
        for (size_t off = 1; off < 70 && off <= in.size(); off += 7) {
            unsigned long long head = 0;
            for (size_t i = 0; i < off; ++i) head += in[i] == '\n';
            if (head + sabreur::detail::host_count_newlines(in.data() + off, in.size() - off) != c) {
                fprintf(stderr, "count differs at offset %zu\n", off);
                return 1;
            }
        }
        printf("%llu\n", c);
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "cut")) {
        printf("%llu\n", (unsigned long long)sabreur::detail::host_guess_cut(in.data(), in.size(), (uint32_t)atoi(argv[2]),
                                                                           (uint64_t)atoll(argv[3])));
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "canon")) {
        std::vector<uint8_t> out;
        sabreur::detail::host_canonical_record(in.data(), in.size(), (uint32_t)atoi(argv[2]), out);
        fwrite(out.data(), 1, out.size(), stdout);
        return 0;
    }
    fprintf(stderr, "usage: host_selftest count | cut FMT WINDOW | canon FMT   (input on stdin)\n");
    return 2;
}
