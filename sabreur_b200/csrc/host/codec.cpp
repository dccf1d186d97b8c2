#include "codec.hpp"
#include "codec_par.hpp"

#include <dlfcn.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <thread>

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
    int (*BZ2_bzBuffToBuffDecompress)(char*, unsigned*, char*, unsigned, int, int) = nullptr;
    // lzma
    int (*lzma_stream_decoder)(lzma_stream*, uint64_t, uint32_t) = nullptr;
    int (*lzma_code)(lzma_stream*, int) = nullptr;
    void (*lzma_end)(lzma_stream*) = nullptr;
    int (*lzma_easy_buffer_encode)(uint32_t, int, const void*, const uint8_t*, size_t, uint8_t*, size_t*, size_t) = nullptr;
    size_t (*lzma_stream_buffer_bound)(size_t) = nullptr;
    int (*lzma_stream_buffer_decode)(uint64_t*, uint32_t, const void*, const uint8_t*, size_t*, size_t, uint8_t*, size_t*, size_t) = nullptr;
    // zstd
    void* (*ZSTD_createDStream)() = nullptr;
    size_t (*ZSTD_freeDStream)(void*) = nullptr;
    size_t (*ZSTD_decompressStream)(void*, ZSTD_outBuffer*, ZSTD_inBuffer*) = nullptr;
    size_t (*ZSTD_compress)(void*, size_t, const void*, size_t, int) = nullptr;
    size_t (*ZSTD_compressBound)(size_t) = nullptr;
    unsigned (*ZSTD_isError)(size_t) = nullptr;
    size_t (*ZSTD_findFrameCompressedSize)(const void*, size_t) = nullptr;
    unsigned long long (*ZSTD_getFrameContentSize)(const void*, size_t) = nullptr;
    size_t (*ZSTD_decompress)(void*, size_t, const void*, size_t) = nullptr;
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
            sym(h, "BZ2_bzBuffToBuffDecompress", L.BZ2_bzBuffToBuffDecompress);
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
            sym(h, "lzma_stream_buffer_decode", L.lzma_stream_buffer_decode);
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
            sym(h, "ZSTD_findFrameCompressedSize", L.ZSTD_findFrameCompressedSize);
            sym(h, "ZSTD_getFrameContentSize", L.ZSTD_getFrameContentSize);
            sym(h, "ZSTD_decompress", L.ZSTD_decompress);
        });
    return L;
}

}  // namespace

// ---- BGZF (blocked gzip, what bgzip / htslib write): every member is a block of <= 64 KiB whose header carries its
// own compressed size ('BC' extra subfield), so the blocks of a group can be found without inflating and inflated in
// parallel.  An ordinary gzip stream is sequential by construction and keeps the one-thread path below.
namespace {

struct BgzfBlock {
    size_t coff, clen;   // raw deflate data inside the group's compressed buffer
    size_t ooff;         // where its bytes go in the group's output
    uint32_t isize, crc;
};

// 0 = not a BGZF block header; else the total size of the block.  h holds the first 18 bytes of the member when the
// 'BC' subfield comes first (as every writer emits it); other layouts are scanned from `extra`.
size_t bgzf_block_size(const uint8_t* h, size_t have, size_t* header_len) {
    if (have < 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
    const size_t xlen = h[10] | ((size_t)h[11] << 8);
    if (have < 12 + xlen) return 0;
    for (size_t q = 12; q + 4 <= 12 + xlen;) {
        const size_t slen = h[q + 2] | ((size_t)h[q + 3] << 8);
        if (h[q] == 'B' && h[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) {
            if (h[3] & ~4) return 0;  // a name / comment / header CRC would sit between the extra field and the data
            *header_len = 12 + xlen;
            return (size_t)(h[q + 4] | ((size_t)h[q + 5] << 8)) + 1;
        }
        q += 4 + slen;
    }
    return 0;
}

class BgzfPool {
public:
    explicit BgzfPool(int n) {
        for (int i = 0; i < n; ++i) th_.emplace_back([this] { work(); });
    }
    ~BgzfPool() {
        {
            std::lock_guard<std::mutex> l(m_);
            quit_ = true;
        }
        go_.notify_all();
        for (auto& t : th_) t.join();
    }
    // inflates blocks[0..n) of `comp` into `out`; throws on corrupt data
    void run(const uint8_t* comp, uint8_t* out, const BgzfBlock* blocks, size_t n) {
        std::unique_lock<std::mutex> l(m_);
        comp_ = comp; out_ = out; blocks_ = blocks; n_ = n;
        next_.store(0);
        pending_ = (int)th_.size();
        bad_ = false;
        ++gen_;
        go_.notify_all();
        done_.wait(l, [this] { return pending_ == 0; });
        if (bad_) throw std::runtime_error("corrupt gzip stream");
    }

private:
    void work() {
        z_stream z;
        std::memset(&z, 0, sizeof z);
        const bool ok = inflateInit2(&z, -15) == Z_OK;  // raw deflate: the block header and trailer are checked here
        uint64_t seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> l(m_);
                go_.wait(l, [&] { return quit_ || gen_ != seen; });
                if (quit_) break;
                seen = gen_;
            }
            bool bad = !ok;
            for (size_t i; !bad && (i = next_.fetch_add(1)) < n_;) {
                const BgzfBlock& b = blocks_[i];
                inflateReset(&z);
                z.next_in = const_cast<Bytef*>(comp_ + b.coff); z.avail_in = (uInt)b.clen;
                uint8_t spare;  // an empty block (the EOF marker) still needs somewhere to "write"
                z.next_out = b.isize ? out_ + b.ooff : &spare; z.avail_out = b.isize ? b.isize : 1;
                const int rc = inflate(&z, Z_FINISH);
                if (rc != Z_STREAM_END || z.total_out != b.isize || z.avail_in != 0) bad = true;
                else if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), out_ + b.ooff, b.isize) != b.crc) bad = true;
            }
            std::lock_guard<std::mutex> l(m_);
            if (bad) bad_ = true;
            if (--pending_ == 0) done_.notify_one();
        }
        if (ok) inflateEnd(&z);
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable go_, done_;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool quit_ = false, bad_ = false;
    std::atomic<size_t> next_{0};
    const uint8_t* comp_ = nullptr;
    uint8_t* out_ = nullptr;
    const BgzfBlock* blocks_ = nullptr;
    size_t n_ = 0;
};

int inflate_threads() {
    if (const char* e = getenv("SABREUR_INFLATE_THREADS")) return std::max(1, atoi(e));
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(16u, hw ? hw : 1u));
}

}  // namespace

// ---- Reader ------------------------------------------------------------------------------------------------
struct Reader::Impl {
    // BGZF mode: groups of blocks are read, inflated by the pool and served from `group_out`
    std::unique_ptr<BgzfPool> pool;
    bool bgzf = false;
    std::vector<uint8_t> group_comp, group_out;
    std::vector<BgzfBlock> group_blocks;
    size_t group_pos = 0, group_len = 0;

    // reads the next group of blocks; false = no further BGZF block (end of file, or an ordinary gzip member follows:
    // the file position then sits at that member's first byte)
    bool next_group() {
        constexpr size_t MAX_BLOCKS = 256;
        group_blocks.clear();
        group_comp.clear();
        size_t out = 0;
        while (group_blocks.size() < MAX_BLOCKS) {
            uint8_t h[18];
            const long at = ftell(f);
            const size_t got = fread(h, 1, sizeof h, f);
            if (got == 0) break;  // end of file
            size_t hlen = 0, total = 0;
            std::vector<uint8_t> ext;
            if (got == sizeof h) {
                total = bgzf_block_size(h, got, &hlen);
                if (!total && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4)) {  // 'BC' is not the first subfield
                    const size_t xlen = h[10] | ((size_t)h[11] << 8);
                    ext.assign(h, h + sizeof h);
                    ext.resize(12 + xlen < sizeof h ? sizeof h : 12 + xlen);
                    if (ext.size() > sizeof h && fread(ext.data() + sizeof h, 1, ext.size() - sizeof h, f) != ext.size() - sizeof h)
                        throw std::runtime_error("truncated compressed stream");
                    total = bgzf_block_size(ext.data(), ext.size(), &hlen);
                }
            }
            if (!total || total < hlen + 8) {  // not a BGZF block: hand the rest of the file to the sequential decoder
                fseek(f, at, SEEK_SET);
                if (group_blocks.empty()) return false;
                break;
            }
            const size_t have = ext.empty() ? sizeof h : ext.size();
            const size_t base = group_comp.size();
            group_comp.resize(base + total);
            std::memcpy(group_comp.data() + base, ext.empty() ? h : ext.data(), std::min(have, total));
            if (total > have && fread(group_comp.data() + base + have, 1, total - have, f) != total - have)
                throw std::runtime_error("truncated compressed stream");
            if (total < have) fseek(f, at + (long)total, SEEK_SET);
            const uint8_t* t = group_comp.data() + base + total - 8;
            BgzfBlock b;
            b.coff = base + hlen;
            b.clen = total - hlen - 8;
            b.crc = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
            b.isize = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
            if (b.isize > (1u << 16)) throw std::runtime_error("corrupt gzip stream");  // BGZF blocks hold <= 64 KiB
            b.ooff = out;
            out += b.isize;
            group_blocks.push_back(b);
        }
        if (group_blocks.empty()) return false;
        group_out.resize(out);
        pool->run(group_comp.data(), group_out.data(), group_blocks.data(), group_blocks.size());
        group_pos = 0;
        group_len = out;
        return true;
    }

    // zstd frames / xz blocks / bzip2 blocks decoded side by side (codec_par.cpp); served unit by unit
    std::unique_ptr<ParallelSource> par;
    ParCodecs par_codecs;
    std::vector<std::vector<uint8_t>> par_out;
    size_t par_unit = 0, par_pos = 0;

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
    if (fmt_ == Format::Gzip) {
        uint8_t h[18];
        size_t hlen = 0;
        const size_t got = fread(h, 1, sizeof h, p_->f);
        fseek(p_->f, 0, SEEK_SET);
        if (bgzf_block_size(h, got, &hlen) && inflate_threads() > 1) {
            p_->bgzf = true;
            p_->pool.reset(new BgzfPool(inflate_threads()));
        }
    }
    if ((fmt_ == Format::Zstd || fmt_ == Format::Lzma || fmt_ == Format::Bzip) && inflate_threads() > 1) {
        const Libs& L = libs(fmt_);
        ParCodecs& c = p_->par_codecs;
        c.ZSTD_findFrameCompressedSize = L.ZSTD_findFrameCompressedSize; c.ZSTD_getFrameContentSize = L.ZSTD_getFrameContentSize;
        c.ZSTD_decompress = L.ZSTD_decompress; c.ZSTD_isError = L.ZSTD_isError;
        c.lzma_stream_buffer_decode = L.lzma_stream_buffer_decode;
        c.BZ2_bzBuffToBuffDecompress = L.BZ2_bzBuffToBuffDecompress;
        p_->par = open_parallel(fmt_, path, c, inflate_threads());  // nullptr: the file does not split, stream it
    }
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
    while (I.bgzf && got < n) {
        if (I.group_pos == I.group_len && !I.next_group()) {
            I.bgzf = false;  // end of file, or ordinary gzip members from here on (the loop below takes them)
            break;
        }
        const size_t take = std::min(n - got, I.group_len - I.group_pos);
        std::memcpy(dst + got, I.group_out.data() + I.group_pos, take);
        I.group_pos += take;
        got += take;
    }
    while (I.par && got < n) {
        if (I.par_unit == I.par_out.size()) {
            I.par_out.clear();
            I.par_unit = I.par_pos = 0;
            if (!I.par->next_group(I.par_out)) {  // the rest of the file (if any) is the sequential decoder's
                const long at = (long)I.par->resume_offset();
                I.par.reset();
                fseek(I.f, 0, SEEK_END);
                if (at >= ftell(I.f)) {  // nothing left: the streaming decoders never see a byte
                    I.file_eof = true;
                    I.stream_end = true;
                    if (I.xz_open) { libs(Format::Lzma).lzma_end(&I.xz); I.xz_open = false; }
                }
                fseek(I.f, at, SEEK_SET);
                break;
            }
            continue;
        }
        const std::vector<uint8_t>& u = I.par_out[I.par_unit];
        const size_t take = std::min(n - got, u.size() - I.par_pos);
        if (take) std::memcpy(dst + got, u.data() + I.par_pos, take);
        I.par_pos += take;
        got += take;
        if (I.par_pos == u.size()) { ++I.par_unit; I.par_pos = 0; }
    }
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
