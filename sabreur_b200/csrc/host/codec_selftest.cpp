// codec_selftest -- CPU-only check of the host codecs: `codec_selftest cat <file>` writes the decompressed stream
// to stdout; `codec_selftest pack <fmt> <level> <member bytes>` compresses stdin member by member to stdout.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "codec.hpp"

using namespace sabreur;

int main(int argc, char** argv) {
    try {
        if (argc >= 3 && !strcmp(argv[1], "cat")) {
            Reader r(argv[2]);
            std::vector<uint8_t> buf(1 << 16);
            size_t chunk = argc > 3 ? (size_t)atol(argv[3]) : buf.size();
            for (size_t n; (n = r.read(buf.data(), std::min(chunk, buf.size()))) > 0;) fwrite(buf.data(), 1, n, stdout);
            return 0;
        }
        if (argc >= 5 && !strcmp(argv[1], "pack")) {
            std::string f = argv[2];
            Format fmt = f == "gz" ? Format::Gzip : f == "bz2" ? Format::Bzip : f == "xz" ? Format::Lzma : f == "zst" ? Format::Zstd : Format::No;
            size_t member = (size_t)atol(argv[4]);
            std::vector<uint8_t> in(member), out;
            for (size_t n; (n = fread(in.data(), 1, member, stdin)) > 0;) {
                compress_member(fmt, atoi(argv[3]), in.data(), n, out);
                fwrite(out.data(), 1, out.size(), stdout);
            }
            return 0;
        }
        if (argc >= 3 && !strcmp(argv[1], "sniff")) {
            printf("%s\n", to_compression_ext(which_format(argv[2])));
            return 0;
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    fprintf(stderr, "usage: codec_selftest cat <file> [chunk] | pack <gz|bz2|xz|zst|no> <level> <member bytes> | sniff <file>\n");
    return 2;
}
