#include "codec.hpp"

#include <dlfcn.h>
#include <zlib.h>

#include <cstring>
#include <mutex>
#include <stdexcept>

namespace sabreur {

const char* to_compression_ext(Format f) {
    switch (f) {
        case Format::Gzip: return ".gz";
        case Format::Bzip: return ".bz2";
        case Format::Lzma: return ".xz";
        case Format::Zstd: return ".zst";
        default: return "";
    }
}

int to_level(int raw) { return (raw >= 1 && raw <= 9) ? raw : 1; }

Format which_format(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open '" + path + "'");
    unsigned char h[5];
    size_t n = fread(h, 1, 5, f);
    fclose(f);
    if (n < 5) throw std::runtime_error("File is too short, less than five bytes");  // niffler::Error::FileTooShort
    if (h[0] == 0x1f && h[1] == 0x8b) return Format::Gzip;
    if (h[0] == 0x42 && h[1] == 0x5a) return Format::Bzip;
    if (h[0] == 0xfd && h[1] == 0x37 && h[2] == 0x7a && h[3] == 0x58 && h[4] == 0x5a) return Format::Lzma;
    if (h[0] == 0x28 && h[1] == 0xb5 && h[2] == 0x2f && h[3] == 0xfd) return Format::Zstd;
    return Format::No;
}

// ---- dlopen'd codecs (prototypes restated from the libraries' public headers) ------------------------
namespace {

struct bz_stream {
    char* next_in; unsigned avail_in; unsigned total_in_lo32; unsigned total_in_hi32;
    char* next_out; unsigned avail_out; unsigned total_out_lo32; unsigned total_out_hi32;
    void* state; void* (*bzalloc)(void*, int, int); void (*bzfree)(void*, void*); void* opaque;
};
constexpr int BZ_OK = 0, BZ_STREAM_END = 4;

struct lzma_stream {
    const uint8_t* next_in; size_t avail_in; uint64_t total_in;
    uint8_t* next_out; size_t avail_out; uint64_t total_out;
    const void* allocator; void* internal;
    void* reserved_ptr1; void* reserved_ptr2; void* reserved_ptr3; void* reserved_ptr4;
    uint64_t reserved_int1; uint64_t reserved_int2; size_t reserved_int3; size_t reserved_int4;
    int reserved_enum1; int reserved_enum2;
};
constexpr int LZMA_OK = 0, LZMA_STREAM_END = 1, LZMA_RUN = 0, LZMA_FINISH = 3, LZMA_CHECK_CRC64 = 4;
constexpr uint32_t LZMA_CONCATENATED = 0x08;

struct ZSTD_inBuffer { const void* src; size_t size; size_t pos; };
struct ZSTD_outBuffer { void* dst; size_t size; size_t pos; };

struct Libs {
    // bz2
    int (*BZ2_bzDecompressInit)(bz_stream*, int, int) = nullptr;
    int (*BZ2_bzDecompress)(bz_stream*) = nullptr;
    int (*BZ2_bzDecompressEnd)(bz_stream*) = nullptr;
    int (*BZ2_bzBuffToBuffCompress)(char*, unsigned*, char*, unsigned, int, int, int) = nullptr;
    // lzma
    int (*lzma_stream_decoder)(lzma_stream*, uint64_t, uint32_t) = nullptr;
    int (*lzma_code)(lzma_stream*, int) = nullptr;
    void (*lzma_end)(lzma_stream*) = nullptr;
    int (*lzma_easy_buffer_encode)(uint32_t, int, const void*, const uint8_t*, size_t, uint8_t*, size_t*, size_t) = nullptr;
    size_t (*lzma_stream_buffer_bound)(size_t) = nullptr;
    // zstd
    void* (*ZSTD_createDStream)() = nullptr;
    size_t (*ZSTD_freeDStream)(void*) = nullptr;
    size_t (*ZSTD_decompressStream)(void*, ZSTD_outBuffer*, ZSTD_inBuffer*) = nullptr;
    size_t (*ZSTD_compress)(void*, size_t, const void*, size_t, int) = nullptr;
    size_t (*ZSTD_compressBound)(size_t) = nullptr;
    unsigned (*ZSTD_isError)(size_t) = nullptr;
};

template <class F>
void sym(void* h, const char* name, F& out) {
    out = reinterpret_cast<F>(dlsym(h, name));
    if (!out) throw std::runtime_error(std::string("missing symbol ") + name);
}

const Libs& libs(Format need) {
    static Libs L;
    static std::once_flag bz, xz, zs;
    if (need == Format::Bzip)
        std::call_once(bz, [] {
            void* h = dlopen("libbz2.so.1.0", RTLD_NOW);
            if (!h) h = dlopen("libbz2.so.1", RTLD_NOW);
            if (!h) throw std::runtime_error("libbz2 is not available on this host");
            sym(h, "BZ2_bzDecompressInit", L.BZ2_bzDecompressInit);
            sym(h, "BZ2_bzDecompress", L.BZ2_bzDecompress);
            sym(h, "BZ2_bzDecompressEnd", L.BZ2_bzDecompressEnd);
            sym(h, "BZ2_bzBuffToBuffCompress", L.BZ2_bzBuffToBuffCompress);
        });
    if (need == Format::Lzma)
        std::call_once(xz, [] {
            void* h = dlopen("liblzma.so.5", RTLD_NOW);
            if (!h) throw std::runtime_error("liblzma is not available on this host");
            sym(h, "lzma_stream_decoder", L.lzma_stream_decoder);
            sym(h, "lzma_code", L.lzma_code);
            sym(h, "lzma_end", L.lzma_end);
            sym(h, "lzma_easy_buffer_encode", L.lzma_easy_buffer_encode);
            sym(h, "lzma_stream_buffer_bound", L.lzma_stream_buffer_bound);
        });
    if (need == Format::Zstd)
        std::call_once(zs, [] {
            void* h = dlopen("libzstd.so.1", RTLD_NOW);
            if (!h) throw std::runtime_error("libzstd is not available on this host");
            sym(h, "ZSTD_createDStream", L.ZSTD_createDStream);
            sym(h, "ZSTD_freeDStream", L.ZSTD_freeDStream);
            sym(h, "ZSTD_decompressStream", L.ZSTD_decompressStream);
            sym(h, "ZSTD_compress", L.ZSTD_compress);
            sym(h, "ZSTD_compressBound", L.ZSTD_compressBound);
            sym(h, "ZSTD_isError", L.ZSTD_isError);
        });
    return L;
}

}  // namespace

// ---- Reader ------------------------------------------------------------------------------------------------
struct Reader::Impl {
    FILE* f = nullptr;
    std::vector<uint8_t> in;
    size_t in_pos = 0, in_len = 0;
    bool file_eof = false, stream_end = true;  // stream_end: the decoder sits between members
    z_stream z{};
    bz_stream bz{};
    lzma_stream xz{};
    void* zd = nullptr;
    bool z_open = false, bz_open = false, xz_open = false;

    bool refill() {
        if (in_pos < in_len) return true;
        if (file_eof) return false;
        in_len = fread(in.data(), 1, in.size(), f);
        in_pos = 0;
        if (in_len == 0) { file_eof = true; return false; }
        return true;
    }
};

Reader::Reader(const std::string& path) : p_(new Impl), fmt_(which_format(path)) {
    p_->f = fopen(path.c_str(), "rb");
    if (!p_->f) throw std::runtime_error("cannot open '" + path + "'");
    p_->in.resize(1 << 20);
    if (fmt_ == Format::Zstd) {
        p_->zd = libs(fmt_).ZSTD_createDStream();
        if (!p_->zd) throw std::runtime_error("ZSTD_createDStream failed");
    } else if (fmt_ == Format::Lzma) {
        std::memset(&p_->xz, 0, sizeof p_->xz);
        if (libs(fmt_).lzma_stream_decoder(&p_->xz, UINT64_MAX, LZMA_CONCATENATED) != LZMA_OK)
            throw std::runtime_error("lzma_stream_decoder failed");
        p_->xz_open = true;
    } else if (fmt_ == Format::Bzip) {
        (void)libs(fmt_);
    }
}

Reader::~Reader() {
    if (p_->z_open) inflateEnd(&p_->z);
    if (p_->bz_open) libs(Format::Bzip).BZ2_bzDecompressEnd(&p_->bz);
    if (p_->xz_open) libs(Format::Lzma).lzma_end(&p_->xz);
    if (p_->zd) libs(Format::Zstd).ZSTD_freeDStream(p_->zd);
    if (p_->f) fclose(p_->f);
}

size_t Reader::read(uint8_t* dst, size_t n) {
    Impl& I = *p_;
    if (n == 0) return 0;
    if (fmt_ == Format::No) return fread(dst, 1, n, I.f);
    size_t got = 0;
    while (got < n) {
        if (!I.refill()) {
            if (fmt_ == Format::Lzma && I.xz_open) {  // flush the concatenated-stream decoder
                I.xz.next_in = nullptr; I.xz.avail_in = 0;
                I.xz.next_out = dst + got; I.xz.avail_out = n - got;
                int rc = libs(fmt_).lzma_code(&I.xz, LZMA_FINISH);
                got = n - I.xz.avail_out;
                if (rc == LZMA_STREAM_END) { libs(fmt_).lzma_end(&I.xz); I.xz_open = false; }
                else if (rc != LZMA_OK) throw std::runtime_error("corrupt xz stream");
                if (rc == LZMA_OK && got < n && I.xz.avail_out == n - got) throw std::runtime_error("truncated xz stream");
                if (!I.xz_open) break;
                continue;
            }
            if (!I.stream_end && fmt_ != Format::Lzma) throw std::runtime_error("truncated compressed stream");
            break;
        }
        if (fmt_ == Format::Gzip) {
            if (I.stream_end) {  // next gzip member
                if (I.z_open) inflateEnd(&I.z);
                std::memset(&I.z, 0, sizeof I.z);
                if (inflateInit2(&I.z, 15 + 16) != Z_OK) throw std::runtime_error("inflateInit2 failed");
                I.z_open = true;
                I.stream_end = false;
            }
            I.z.next_in = I.in.data() + I.in_pos; I.z.avail_in = (uInt)(I.in_len - I.in_pos);
            I.z.next_out = dst + got; I.z.avail_out = (uInt)std::min<size_t>(n - got, 1u << 30);
            const uInt out0 = I.z.avail_out;
            int rc = inflate(&I.z, Z_NO_FLUSH);
            I.in_pos = I.in_len - I.z.avail_in;
            got += out0 - I.z.avail_out;
            if (rc == Z_STREAM_END) I.stream_end = true;
            else if (rc != Z_OK && rc != Z_BUF_ERROR) throw std::runtime_error("corrupt gzip stream");
        } else if (fmt_ == Format::Bzip) {
            const Libs& L = libs(fmt_);
            if (I.stream_end) {
                if (I.bz_open) L.BZ2_bzDecompressEnd(&I.bz);
                std::memset(&I.bz, 0, sizeof I.bz);
                if (L.BZ2_bzDecompressInit(&I.bz, 0, 0) != BZ_OK) throw std::runtime_error("BZ2_bzDecompressInit failed");
                I.bz_open = true;
                I.stream_end = false;
            }
            I.bz.next_in = (char*)I.in.data() + I.in_pos; I.bz.avail_in = (unsigned)(I.in_len - I.in_pos);
            I.bz.next_out = (char*)dst + got; I.bz.avail_out = (unsigned)std::min<size_t>(n - got, 1u << 30);
            const unsigned out0 = I.bz.avail_out;
            int rc = L.BZ2_bzDecompress(&I.bz);
            I.in_pos = I.in_len - I.bz.avail_in;
            got += out0 - I.bz.avail_out;
            if (rc == BZ_STREAM_END) I.stream_end = true;
            else if (rc != BZ_OK) throw std::runtime_error("corrupt bzip2 stream");
        } else if (fmt_ == Format::Lzma) {
            const Libs& L = libs(fmt_);
            I.xz.next_in = I.in.data() + I.in_pos; I.xz.avail_in = I.in_len - I.in_pos;
            I.xz.next_out = dst + got; I.xz.avail_out = n - got;
            int rc = L.lzma_code(&I.xz, LZMA_RUN);
            I.in_pos = I.in_len - I.xz.avail_in;
            got = n - I.xz.avail_out;
            if (rc == LZMA_STREAM_END) { L.lzma_end(&I.xz); I.xz_open = false; I.file_eof = true; I.in_pos = I.in_len; break; }
            if (rc != LZMA_OK) throw std::runtime_error("corrupt xz stream");
        } else {  // zstd
            const Libs& L = libs(fmt_);
            ZSTD_inBuffer ib{I.in.data(), I.in_len, I.in_pos};
            ZSTD_outBuffer ob{dst, n, got};
            size_t rc = L.ZSTD_decompressStream(I.zd, &ob, &ib);
            if (L.ZSTD_isError(rc)) throw std::runtime_error("corrupt zstd stream");
            I.in_pos = ib.pos;
            got = ob.pos;
            I.stream_end = (rc == 0);  // a frame was completely decoded and flushed
        }
    }
    return got;
}

// ---- member-at-a-time compression ------------------------------------------------------------------------------
void compress_member(Format f, int level, const uint8_t* in, size_t n, std::vector<uint8_t>& out) {
    out.clear();
    if (n == 0) return;
    level = to_level(level);
    switch (f) {
        case Format::No:
            out.assign(in, in + n);
            return;
        case Format::Gzip: {
            z_stream z{};
            if (deflateInit2(&z, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK)
                throw std::runtime_error("deflateInit2 failed");
            out.resize(deflateBound(&z, (uLong)n) + 32);
            z.next_in = const_cast<Bytef*>(in); z.avail_in = (uInt)n;  // callers keep members below 1 GiB
            z.next_out = out.data(); z.avail_out = (uInt)out.size();
            int rc = deflate(&z, Z_FINISH);
            size_t produced = out.size() - z.avail_out;
            deflateEnd(&z);
            if (rc != Z_STREAM_END) throw std::runtime_error("deflate failed");
            out.resize(produced);
            return;
        }
        case Format::Bzip: {
            unsigned cap = (unsigned)(n + n / 100 + 600 + 64);
            out.resize(cap);
            int rc = libs(f).BZ2_bzBuffToBuffCompress((char*)out.data(), &cap, (char*)const_cast<uint8_t*>(in), (unsigned)n, level, 0, 0);
            if (rc != BZ_OK) throw std::runtime_error("BZ2_bzBuffToBuffCompress failed");
            out.resize(cap);
            return;
        }
        case Format::Lzma: {
            const Libs& L = libs(f);
            out.resize(L.lzma_stream_buffer_bound(n));
            size_t pos = 0;
            int rc = L.lzma_easy_buffer_encode((uint32_t)level, LZMA_CHECK_CRC64, nullptr, in, n, out.data(), &pos, out.size());
            if (rc != LZMA_OK) throw std::runtime_error("lzma_easy_buffer_encode failed");
            out.resize(pos);
            return;
        }
        case Format::Zstd: {
            const Libs& L = libs(f);
            out.resize(L.ZSTD_compressBound(n));
            size_t rc = L.ZSTD_compress(out.data(), out.size(), in, n, level);
            if (L.ZSTD_isError(rc)) throw std::runtime_error("ZSTD_compress failed");
            out.resize(rc);
            return;
        }
    }
}

}  // namespace sabreur
