// codec_par.cpp -- block-parallel decompression of the input formats that can be split WITHOUT decoding
// ("parallel decompression feeding pinned buffers"; the reference reads its inputs through
// niffler::send::from_path on one thread, src/demux.rs:22, 81-82).
//
//   zstd   a file is a sequence of frames; ZSTD_findFrameCompressedSize walks a frame's block headers (no decoding)
//          and the frame header usually carries the decompressed size: frames are decoded side by side.  A frame
//          without a content size, or a giant single frame, hands over to the streaming decoder.
//   xz     every stream ends in an index of its blocks (unpadded + uncompressed sizes) and a footer pointing at the
//          index: read from the END of the file backwards, that gives every stream and every block without decoding.
//          Each block is wrapped into a tiny single-block stream of its own (copied header, fresh index + footer) and
//          handed to lzma_stream_buffer_decode, so blocks of one stream (xz -T) and concatenated streams (what this
//          host writes) both decode in parallel.
//   bzip2  blocks start with the 48-bit magic 0x314159265359 at ANY bit offset and carry their own CRC; a stream ends
//          with 0x177245385090 + the combined CRC.  The compressed bits are scanned for the magics, every block is
//          shifted to a byte boundary and wrapped as a one-block stream ("BZh<level>" + block + end magic + block
//          CRC), and the library decodes the wrapped blocks in parallel -- the lbzip2 / pbzip2 technique.  A magic
//          that turns out to be payload (2^-48 per position) makes a wrapped block fail: it is then glued to its
//          successor and tried again.
//   gzip   deflate has no block index: a plain .gz stays sequential by format (BGZF, whose members carry their own
//          size, is handled in codec.cpp).
//
// Everything here only finds and decodes units; codec.cpp's Reader serves their bytes in order.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "codec_par.hpp"

namespace sabreur {

namespace {

// runs body(i) for i in [0, n) on up to `threads` threads; the first exception is rethrown on the caller
template <class F>
void parallel_units(size_t n, int threads, F&& body) {
    std::atomic<size_t> next{0};
    std::exception_ptr err;
    std::mutex em;
    auto work = [&] {
        try {
            for (size_t i; (i = next.fetch_add(1)) < n;) body(i);
        } catch (...) {
            std::lock_guard<std::mutex> l(em);
            if (!err) err = std::current_exception();
            next.store(n);
        }
    };
    const int nt = (int)std::min<size_t>((size_t)std::max(1, threads), n);
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (err) std::rethrow_exception(err);
}

struct Mapped {  // the compressed file, read-only
    const uint8_t* p = nullptr;
    size_t n = 0;
    ~Mapped() {
        if (p && n) munmap(const_cast<uint8_t*>(p), n);
    }
    bool open(const std::string& path) {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size <= 0) { close(fd); return false; }
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (m == MAP_FAILED) return false;
        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
        p = (const uint8_t*)m;
        n = (size_t)st.st_size;
        return true;
    }
};

constexpr size_t GROUP_OUT_BYTES = 256u << 20;  // decoded bytes a group may hold (plus one unit)

// ------------------------------------------------------------------------------------------------------------
// zstd: frame-parallel
// ------------------------------------------------------------------------------------------------------------
class ZstdFrames : public ParallelSource {
public:
    ZstdFrames(const ParCodecs& c, int threads) : c_(c), threads_(threads) {}
    bool open(const std::string& path) {
        if (!c_.ZSTD_findFrameCompressedSize || !c_.ZSTD_getFrameContentSize || !c_.ZSTD_decompress) return false;
        if (!m_.open(path)) return false;
        // worth it only when the file has more than one frame with a known size
        size_t sz, cs;
        if (!frame_at(0, sz, cs) || sz >= m_.n) return false;
        return true;
    }
    bool next_group(std::vector<std::vector<uint8_t>>& outs) override {
        struct U { size_t off, sz, cs; };
        std::vector<U> units;
        size_t total = 0;
        while (pos_ < m_.n && total < GROUP_OUT_BYTES && units.size() < 4096) {
            size_t sz, cs;
            if (!frame_at(pos_, sz, cs)) break;  // the streaming decoder takes over at pos_
            units.push_back(U{pos_, sz, cs});
            pos_ += sz;
            total += cs;
        }
        if (units.empty()) return false;
        outs.assign(units.size(), {});
        parallel_units(units.size(), threads_, [&](size_t i) {
            outs[i].resize(units[i].cs);
            const size_t rc = c_.ZSTD_decompress(outs[i].data(), units[i].cs, m_.p + units[i].off, units[i].sz);
            if (c_.ZSTD_isError(rc) || rc != units[i].cs) throw std::runtime_error("corrupt zstd stream");
        });
        return true;
    }
    size_t resume_offset() const override { return pos_; }

private:
    bool frame_at(size_t at, size_t& sz, size_t& cs) const {
        if (m_.n - at < 8) return false;
        sz = c_.ZSTD_findFrameCompressedSize(m_.p + at, m_.n - at);
        if (c_.ZSTD_isError(sz) || sz == 0 || sz > m_.n - at) return false;
        const unsigned long long c = c_.ZSTD_getFrameContentSize(m_.p + at, m_.n - at);
        if (c >= 0xFFFFFFFFFFFFFFFEull) return false;   // unknown (-1) or error (-2)
        if (c > (1ull << 30)) return false;             // a giant frame: stream it instead of buffering it
        cs = (size_t)c;
        return true;
    }
    const ParCodecs& c_;
    int threads_;
    Mapped m_;
    size_t pos_ = 0;
};

// ------------------------------------------------------------------------------------------------------------
// xz: block-parallel from the stream indexes
// ------------------------------------------------------------------------------------------------------------
uint32_t le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
uint32_t crc_of(const uint8_t* p, size_t n) { return (uint32_t)crc32(crc32(0L, Z_NULL, 0), p, (uInt)n); }

bool get_vli(const uint8_t* p, size_t n, size_t& at, uint64_t& v) {  // xz variable-length integer, <= 9 bytes
    v = 0;
    for (int i = 0; i < 9; ++i) {
        if (at >= n) return false;
        const uint8_t b = p[at++];
        v |= (uint64_t)(b & 0x7F) << (7 * i);
        if (!(b & 0x80)) return !(b == 0 && i > 0);
    }
    return false;
}
void put_vli(std::vector<uint8_t>& o, uint64_t v) {
    while (v >= 0x80) { o.push_back((uint8_t)(v | 0x80)); v >>= 7; }
    o.push_back((uint8_t)v);
}

class XzBlocks : public ParallelSource {
public:
    XzBlocks(const ParCodecs& c, int threads) : c_(c), threads_(threads) {}
    bool open(const std::string& path) {
        if (!c_.lzma_stream_buffer_decode || !m_.open(path)) return false;
        if (!parse()) return false;
        return units_.size() > 1;  // one block: nothing to run side by side
    }
    bool next_group(std::vector<std::vector<uint8_t>>& outs) override {
        const size_t first = next_;
        size_t total = 0;
        while (next_ < units_.size() && total < GROUP_OUT_BYTES) total += units_[next_++].uncomp;
        if (next_ == first) return false;
        outs.assign(next_ - first, {});
        parallel_units(next_ - first, threads_, [&](size_t i) { decode(units_[first + i], outs[i]); });
        return true;
    }
    size_t resume_offset() const override { return m_.n; }  // the whole file is covered (or open() refused it)

private:
    struct Unit { size_t off; uint64_t unpadded, uncomp; uint8_t flags[2]; };

    // all streams of the file, last to first; false = something this parser does not understand (sequential path)
    bool parse() {
        size_t pos = m_.n;
        std::vector<Unit> rev;
        while (pos > 0) {
            while (pos >= 4 && le32(m_.p + pos - 4) == 0) pos -= 4;  // stream padding
            if (pos == 0) break;
            if (pos < 32 || (pos & 3)) return false;
            const uint8_t* f = m_.p + pos - 12;  // footer: crc32 | backward size | flags | "YZ"
            if (f[10] != 'Y' || f[11] != 'Z' || le32(f) != crc_of(f + 4, 6)) return false;
            const size_t index_size = ((size_t)le32(f + 4) + 1) * 4;
            if (index_size + 24 > pos) return false;
            const size_t index_at = pos - 12 - index_size;
            const uint8_t* ix = m_.p + index_at;
            if (ix[0] != 0 || le32(ix + index_size - 4) != crc_of(ix, index_size - 4)) return false;
            size_t at = 1;
            uint64_t nrec;
            if (!get_vli(ix, index_size - 4, at, nrec) || nrec > index_size / 2) return false;  // (a record takes >= 2 bytes)
            std::vector<Unit> blocks((size_t)nrec);
            uint64_t blocks_bytes = 0;
            for (auto& b : blocks) {
                if (!get_vli(ix, index_size - 4, at, b.unpadded) || !get_vli(ix, index_size - 4, at, b.uncomp)) return false;
                if (b.unpadded < 5 || b.uncomp > (1ull << 32)) return false;
                b.flags[0] = f[8];
                b.flags[1] = f[9];
                blocks_bytes += (b.unpadded + 3) & ~3ull;
            }
            if (blocks_bytes + 12 > index_at) return false;
            const size_t stream_at = index_at - (size_t)blocks_bytes - 12;
            const uint8_t* h = m_.p + stream_at;  // header: FD 37 7A 58 5A 00 | flags | crc32
            static const uint8_t MAGIC[6] = {0xFD, 0x37, 0x7A, 0x58, 0x5A, 0x00};
            if (memcmp(h, MAGIC, 6) != 0 || h[6] != f[8] || h[7] != f[9] || le32(h + 8) != crc_of(h + 6, 2)) return false;
            size_t off = stream_at + 12;
            for (auto& b : blocks) { b.off = off; off += (size_t)((b.unpadded + 3) & ~3ull); }
            for (size_t i = blocks.size(); i-- > 0;) rev.push_back(blocks[i]);
            pos = stream_at;
        }
        units_.assign(rev.rbegin(), rev.rend());
        return !units_.empty();
    }

    // the block wrapped as a stream of its own: header | block (+ padding) | index of one record | footer
    void decode(const Unit& u, std::vector<uint8_t>& out) const {
        const size_t padded = (size_t)((u.unpadded + 3) & ~3ull);
        std::vector<uint8_t> s;
        s.reserve(12 + padded + 40);
        static const uint8_t MAGIC[6] = {0xFD, 0x37, 0x7A, 0x58, 0x5A, 0x00};
        s.insert(s.end(), MAGIC, MAGIC + 6);
        s.push_back(u.flags[0]);
        s.push_back(u.flags[1]);
        s.resize(12);
        put32(s.data() + 8, crc_of(s.data() + 6, 2));
        s.insert(s.end(), m_.p + u.off, m_.p + u.off + padded);
        const size_t index_at = s.size();
        s.push_back(0);
        put_vli(s, 1);
        put_vli(s, u.unpadded);
        put_vli(s, u.uncomp);
        while ((s.size() - index_at) & 3) s.push_back(0);
        s.resize(s.size() + 4);
        put32(s.data() + s.size() - 4, crc_of(s.data() + index_at, s.size() - 4 - index_at));
        const size_t index_size = s.size() - index_at;
        s.resize(s.size() + 12);
        uint8_t* f = s.data() + s.size() - 12;
        put32(f + 4, (uint32_t)(index_size / 4 - 1));
        f[8] = u.flags[0];
        f[9] = u.flags[1];
        f[10] = 'Y';
        f[11] = 'Z';
        put32(f, crc_of(f + 4, 6));
        out.resize((size_t)u.uncomp);
        uint64_t memlimit = UINT64_MAX;
        size_t in_pos = 0, out_pos = 0;
        uint8_t spare;
        const int rc = c_.lzma_stream_buffer_decode(&memlimit, 0, nullptr, s.data(), &in_pos, s.size(), out.empty() ? &spare : out.data(),
                                                    &out_pos, out.size());
        if (rc != 0 || out_pos != out.size() || in_pos != s.size()) throw std::runtime_error("corrupt xz stream");
    }

    const ParCodecs& c_;
    int threads_;
    Mapped m_;
    std::vector<Unit> units_;
    size_t next_ = 0;
};

// ------------------------------------------------------------------------------------------------------------
// bzip2: block-parallel by magic scan
// ------------------------------------------------------------------------------------------------------------
constexpr uint64_t BZ_BLOCK = 0x314159265359ull, BZ_EOS = 0x177245385090ull, M48 = (1ull << 48) - 1;

class Bz2Blocks : public ParallelSource {
public:
    Bz2Blocks(const ParCodecs& c, int threads) : c_(c), threads_(threads) {}
    bool open(const std::string& path) {
        if (!c_.BZ2_bzBuffToBuffDecompress || !m_.open(path)) return false;
        return m_.n >= 14 && header_at(0);
    }
    // One group = the blocks found in the next window of compressed bytes.
    bool next_group(std::vector<std::vector<uint8_t>>& outs) override {
        if (done_) return false;
        struct B { uint64_t bit0, bit1; uint8_t level; };  // [bit0, bit1): from the block magic to the next magic
        std::vector<B> blocks;
        const uint64_t total_bits = (uint64_t)m_.n * 8;
        while (blocks.size() < (size_t)threads_ * 4 && !done_) {
            if (!in_stream_) {  // at a stream header (byte aligned) or at the end of the file
                size_t at = (size_t)((bit_ + 7) / 8);
                if (at >= m_.n) { done_ = true; break; }
                if (!header_at(at)) throw std::runtime_error("corrupt bzip2 stream");  // trailing bytes that are no stream
                level_ = m_.p[at + 3];
                bit_ = (uint64_t)(at + 4) * 8;
                in_stream_ = true;
                if (peek48(bit_) == BZ_EOS) {  // an empty stream: header, end magic, CRC
                    bit_ += 80;
                    in_stream_ = false;
                    continue;
                }
                if (peek48(bit_) != BZ_BLOCK) throw std::runtime_error("corrupt bzip2 stream");
            }
            // bit_ sits on a block magic: find the next magic (block or end of stream).  An end-of-stream magic is only
            // believed when what follows it (CRC, padding to a byte) is the end of the file or another stream's header
            // and first magic: 48 payload bits that merely look like it are skipped
            uint64_t q = bit_ + 48 + 32;
            bool eos = false;
            while (true) {
                q = find_magic(q, total_bits, eos);
                if (q == UINT64_MAX) throw std::runtime_error("truncated compressed stream");
                if (!eos) break;
                const size_t at = (size_t)((q + 80 + 7) / 8);
                if (at >= m_.n) break;
                if (header_at(at)) {
                    const uint64_t v = peek48((uint64_t)(at + 4) * 8);
                    if (v == BZ_BLOCK || v == BZ_EOS) break;
                }
                ++q;
            }
            blocks.push_back(B{bit_, q, level_});
            if (eos) {
                bit_ = q + 48 + 32;
                in_stream_ = false;
            } else {
                bit_ = q;
            }
        }
        if (blocks.empty()) return false;
        outs.assign(blocks.size(), {});
        std::vector<uint8_t> ok(blocks.size(), 0);
        parallel_units(blocks.size(), threads_, [&](size_t i) { ok[i] = decode_bits(blocks[i].bit0, blocks[i].bit1, blocks[i].level, outs[i]) ? 1 : 0; });
        // a "magic" inside a block's payload splits it in two pieces that both fail: glue such neighbours and try again
        for (size_t i = 0; i < blocks.size(); ++i) {
            if (ok[i]) continue;
            size_t j = i;
            bool fixed = false;
            while (!fixed && j + 1 < blocks.size() && !ok[j + 1]) {
                ++j;
                fixed = decode_bits(blocks[i].bit0, blocks[j].bit1, blocks[i].level, outs[i]);
            }
            if (!fixed) throw std::runtime_error("corrupt bzip2 stream");
            for (size_t k = i + 1; k <= j; ++k) { outs[k].clear(); ok[k] = 1; }
            ok[i] = 1;
        }
        return true;
    }
    size_t resume_offset() const override { return m_.n; }

private:
    bool header_at(size_t at) const {
        return at + 4 <= m_.n && m_.p[at] == 'B' && m_.p[at + 1] == 'Z' && m_.p[at + 2] == 'h' && m_.p[at + 3] >= '1' && m_.p[at + 3] <= '9';
    }
    // 48 bits starting at bit position `bit` (bit 0 = the most significant bit of byte 0), zero padded past the end
    uint64_t peek48(uint64_t bit) const {
        const size_t by = (size_t)(bit >> 3);
        uint64_t w = 0;
        for (size_t k = 0; k < 8; ++k) w = (w << 8) | (by + k < m_.n ? m_.p[by + k] : 0);
        return (w >> (16 - (bit & 7))) & M48;
    }
    // first bit position >= from where a block or end-of-stream magic starts
    uint64_t find_magic(uint64_t from, uint64_t total_bits, bool& eos) const {
        if (from + 48 > total_bits) return UINT64_MAX;
        size_t by = (size_t)(from >> 3);
        // finish the partial first byte bit by bit, then whole bytes: an 8-byte window per byte, eight shifts each
        for (uint64_t b = from; b < (uint64_t)(by + 1) * 8 && b + 48 <= total_bits; ++b) {
            const uint64_t v = peek48(b);
            if (v == BZ_BLOCK || v == BZ_EOS) { eos = v == BZ_EOS; return b; }
        }
        ++by;
        const size_t last = m_.n >= 8 ? m_.n - 8 : 0;  // windows that lie wholly inside the file
        uint64_t w = 0;
        if (by <= last) {
            for (size_t k = 0; k < 7; ++k) w = (w << 8) | m_.p[by + k];
            for (; by <= last; ++by) {
                w = (w << 8) | m_.p[by + 7];
                // quick reject: both magics, at any of the eight shifts, need the byte pattern of their second byte
                // somewhere in the window -- cheaper to just test the eight shifts with two compares each
                for (int s = 0; s < 8; ++s) {
                    const uint64_t v = (w >> (16 - s)) & M48;
                    if (v == BZ_BLOCK || v == BZ_EOS) { eos = v == BZ_EOS; return (uint64_t)by * 8 + (uint64_t)s; }
                }
            }
        }
        for (uint64_t b = (uint64_t)by * 8; b + 48 <= total_bits; ++b) {  // the file's last few bytes
            const uint64_t v = peek48(b);
            if (v == BZ_BLOCK || v == BZ_EOS) { eos = v == BZ_EOS; return b; }
        }
        return UINT64_MAX;
    }
    // the bits [bit0, bit1) (a block: magic, CRC, payload) as a complete one-block stream -> decoded bytes
    bool decode_bits(uint64_t bit0, uint64_t bit1, uint8_t level, std::vector<uint8_t>& out) const {
        const uint64_t nbits = bit1 - bit0;
        std::vector<uint8_t> s;
        s.reserve((size_t)(nbits / 8) + 32);
        s.push_back('B'); s.push_back('Z'); s.push_back('h'); s.push_back(level);
        const size_t by0 = (size_t)(bit0 >> 3);
        const unsigned sh = (unsigned)(bit0 & 7);
        const size_t whole = (size_t)(nbits / 8);
        for (size_t k = 0; k < whole; ++k) {
            const uint8_t a = m_.p[by0 + k], b = by0 + k + 1 < m_.n ? m_.p[by0 + k + 1] : 0;
            s.push_back((uint8_t)((a << sh) | (sh ? b >> (8 - sh) : 0)));
        }
        // the tail: the block's last (nbits % 8) bits, then the end-of-stream magic and the block's CRC as the stream's
        uint64_t acc = 0;
        unsigned nacc = (unsigned)(nbits & 7);
        if (nacc) {
            const uint8_t a = m_.p[by0 + whole], b = by0 + whole + 1 < m_.n ? m_.p[by0 + whole + 1] : 0;
            const uint8_t v = (uint8_t)((a << sh) | (sh ? b >> (8 - sh) : 0));
            acc = v >> (8 - nacc);
        }
        auto put_bits = [&](uint64_t v, unsigned n) {
            for (unsigned i = n; i-- > 0;) {
                acc = (acc << 1) | ((v >> i) & 1);
                if (++nacc == 8) { s.push_back((uint8_t)acc); acc = 0; nacc = 0; }
            }
        };
        put_bits(BZ_EOS, 48);
        put_bits(peek48(bit0 + 48) >> 16, 32);  // the block CRC follows the block magic
        if (nacc) { s.push_back((uint8_t)(acc << (8 - nacc))); }
        for (size_t cap = (size_t)(level - '0') * 100000 + 4096;; cap *= 8) {
            out.resize(cap);
            unsigned got = (unsigned)cap;
            const int rc = c_.BZ2_bzBuffToBuffDecompress((char*)out.data(), &got, (char*)s.data(), (unsigned)s.size(), 0, 0);
            if (rc == 0) { out.resize(got); return true; }
            if (rc != -8 /* BZ_OUTBUFF_FULL: long runs expand past the block size */ || cap > (1u << 29)) { out.clear(); return false; }
        }
    }

    const ParCodecs& c_;
    int threads_;
    Mapped m_;
    uint64_t bit_ = 0;
    bool in_stream_ = false, done_ = false;
    uint8_t level_ = '9';
};

}  // namespace

std::unique_ptr<ParallelSource> open_parallel(Format f, const std::string& path, const ParCodecs& codecs, int threads) {
    if (threads <= 1) return nullptr;
    if (f == Format::Zstd) {
        std::unique_ptr<ZstdFrames> s(new ZstdFrames(codecs, threads));
        if (s->open(path)) return s;
    } else if (f == Format::Lzma) {
        std::unique_ptr<XzBlocks> s(new XzBlocks(codecs, threads));
        if (s->open(path)) return s;
    } else if (f == Format::Bzip) {
        std::unique_ptr<Bz2Blocks> s(new Bz2Blocks(codecs, threads));
        if (s->open(path)) return s;
    }
    return nullptr;
}

}  // namespace sabreur
