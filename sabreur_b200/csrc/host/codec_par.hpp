// codec_par.hpp -- block-parallel decompression of splittable inputs (zstd frames, xz blocks, bzip2 blocks): see
// codec_par.cpp.  codec.cpp's Reader asks open_parallel() first and serves the decoded units in order; when a file
// cannot be split (one frame / one block, an unknown layout, a pipe) it gets nullptr and streams on one thread.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "codec.hpp"

namespace sabreur {

// the one-shot entry points of the dlopen'd libraries (filled by codec.cpp; a null pointer = not available)
struct ParCodecs {
    size_t (*ZSTD_findFrameCompressedSize)(const void*, size_t) = nullptr;
    unsigned long long (*ZSTD_getFrameContentSize)(const void*, size_t) = nullptr;
    size_t (*ZSTD_decompress)(void*, size_t, const void*, size_t) = nullptr;
    unsigned (*ZSTD_isError)(size_t) = nullptr;
    int (*lzma_stream_buffer_decode)(uint64_t*, uint32_t, const void*, const uint8_t*, size_t*, size_t, uint8_t*, size_t*, size_t) = nullptr;
    int (*BZ2_bzBuffToBuffDecompress)(char*, unsigned*, char*, unsigned, int, int) = nullptr;
};

class ParallelSource {
public:
    virtual ~ParallelSource() {}
    // decodes the next group of independent units on the thread pool: outs[0], outs[1] ... hold the stream's next
    // bytes in order.  false = nothing (more) to decode in parallel; the sequential decoder then continues at
    // resume_offset() (== the file size when the whole file has been served).
    virtual bool next_group(std::vector<std::vector<uint8_t>>& outs) = 0;
    virtual size_t resume_offset() const = 0;
};

std::unique_ptr<ParallelSource> open_parallel(Format f, const std::string& path, const ParCodecs& codecs, int threads);

}  // namespace sabreur
