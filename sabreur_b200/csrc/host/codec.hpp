// codec.hpp -- input decompression and output compression of the sabreur host (the part niffler plays in the
// reference: src/demux.rs:22,81-82 niffler::send::from_path; src/utils.rs:125-130 sniff; utils.rs:139 get_writer).
// zlib is linked; libbz2 / liblzma / libzstd ship without headers in this image and are bound with dlopen.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

namespace sabreur {

enum class Format { No, Gzip, Bzip, Lzma, Zstd };  // niffler::send::compression::Format

const char* to_compression_ext(Format f);           // src/utils.rs:76-84
Format which_format(const std::string& path);       // src/utils.rs:125-130; throws std::runtime_error (< 5 bytes: "file too short")
int to_level(int raw);                              // src/utils.rs:87-100: 1..9, anything else -> 1

// streaming reader of a possibly compressed file (multi-member gzip, concatenated bz2 / xz / zstd streams)
class Reader {
public:
    explicit Reader(const std::string& path);
    ~Reader();
    Reader(const Reader&) = delete;
    Reader& operator=(const Reader&) = delete;
    Format format() const { return fmt_; }
    // reads up to n decompressed bytes; 0 = end of stream.  Throws std::runtime_error on corrupt input.
    size_t read(uint8_t* dst, size_t n);

private:
    struct Impl;
    std::unique_ptr<Impl> p_;
    Format fmt_;
};

// one self-contained compressed member / frame / stream holding `n` bytes; members concatenate into a valid file
// of that format (the reference writes one member per RECORD, utils.rs:139; the decompressed content is identical).
void compress_member(Format f, int level, const uint8_t* in, size_t n, std::vector<uint8_t>& out);

}  // namespace sabreur
