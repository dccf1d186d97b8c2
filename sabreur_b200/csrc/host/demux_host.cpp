// demux_host.cpp -- host pipeline around the C ABI: decompress -> pinned slots -> GPU(s) -> per-barcode writers.
//
//   reader thread   decompresses the input into pinned slot buffers (sbr_slot_buffer) and cuts batches at a GUESSED
//                   record boundary so that several batches can be in flight on several GPUs at once;
//   GPU thread      the caller's thread: submits batches round-robin (batch s -> GPU s mod G), waits for them in
//                   sequence order and VERIFIES every guess against the scan (`consumed`): exact record boundaries
//                   come from the GPU alone.  A wrong guess costs one re-submission of the following batch with the
//                   unconsumed tail in front of it (every slot keeps room there);
//   writer thread   takes results in batch order, gathers every bin's records (order[] / rec_off[]) into one
//                   buffer per output file, compresses it as one member and appends it: files keep input order
//                   exactly like the sequential reference (src/demux.rs:49-66, utils.rs:133-158).
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include "demux_host.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <exception>
#include <mutex>
#include <thread>

#include <unistd.h>

#include "../../../include/sabreur_gpu.h"

namespace sabreur {

std::vector<std::vector<std::string>> split_by_tab(const std::string& content) {
    if (content.find('\t') == std::string::npos) throw std::runtime_error("Input string is not tab-delimited");
    std::vector<std::vector<std::string>> rows;
    size_t pos = 0;
    while (pos < content.size()) {  // str::lines(): split at '\n', a trailing '\r' is dropped, no empty last line
        size_t nl = content.find('\n', pos);
        std::string line = content.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
        pos = nl == std::string::npos ? content.size() : nl + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::vector<std::string> f;
        size_t a = 0;
        while (true) {
            size_t t = line.find('\t', a);
            f.push_back(line.substr(a, t == std::string::npos ? std::string::npos : t - a));
            if (t == std::string::npos) break;
            a = t + 1;
        }
        rows.push_back(std::move(f));
    }
    return rows;
}

namespace {

using Clock = std::chrono::steady_clock;
double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

struct Gpu {
    sbr_ctx* ctx = nullptr;
    std::vector<uint8_t*> slot_host;
};

struct Batch {
    uint64_t seq = 0;
    int gpu = 0;
    uint32_t slot = 0;
    uint64_t off = 0, n = 0;  // the batch is slot_host[off .. off + n)
    bool final_batch = false;   // the stream ends with this batch
    bool exact_end = false;     // the batch ends exactly at a record boundary (FASTA cut in front of a '>' line)
    sbr_result res{};
};

template <class T>
class Channel {  // bounded by construction (every item holds a slot)
public:
    void push(T v) {
        { std::lock_guard<std::mutex> l(m_); q_.push_back(std::move(v)); }
        cv_.notify_all();
    }
    void close() {
        { std::lock_guard<std::mutex> l(m_); closed_ = true; }
        cv_.notify_all();
    }
    bool pop(T& out) {  // false: closed and drained
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        return true;
    }
    bool try_pop(T& out) {
        std::lock_guard<std::mutex> l(m_);
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        return true;
    }
    bool drained() {
        std::lock_guard<std::mutex> l(m_);
        return closed_ && q_.empty();
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> q_;
    bool closed_ = false;
};

// ---- guessed record boundary near the end of [p, p + n): the offset where the last incomplete record starts ----
// FASTQ: a line that starts with '@' and whose second-next line starts with '+' ('@' also opens quality lines, but
// then the second-next line is a sequence).  FASTA: the last line starting with '>'.  Only a guess: the GPU verifies.
uint64_t guess_cut(const uint8_t* p, uint64_t n, uint32_t fmt, uint64_t window) {
    const uint64_t lo = n > window ? n - window : 0;
    std::vector<uint64_t> starts;  // line starts, descending
    uint64_t e = n;
    while (e > lo && starts.size() < 4096) {
        const void* r = memrchr(p + lo, '\n', e - lo);
        if (!r) break;
        uint64_t nl = (const uint8_t*)r - p;
        if (nl + 1 < n) starts.push_back(nl + 1);
        e = nl;
    }
    if (fmt == SBR_FMT_FASTA) {
        for (uint64_t s : starts)
            if (p[s] == '>') return s;
        return n;
    }
    for (size_t j = 2; j < starts.size(); ++j)  // starts[j-2] is two lines after starts[j]
        if (p[starts[j]] == '@' && p[starts[j - 2]] == '+') return starts[j];
    return n;
}

// newlines in [p, p + n): 64 bytes per step with byte-wide counters (SSE2 is part of x86-64), ~30 GB/s per core
static uint64_t count_newlines(const uint8_t* p, uint64_t n) {
    uint64_t c = 0, i = 0;
#if defined(__x86_64__)
    const __m128i nl = _mm_set1_epi8('\n'), zero = _mm_setzero_si128();
    while (n - i >= 64) {
        uint64_t blocks = (n - i) / 64;
        if (blocks > 255) blocks = 255;  // byte counters: at most 255 increments before they are folded
        __m128i a0 = zero, a1 = zero, a2 = zero, a3 = zero;
        for (uint64_t b = 0; b < blocks; ++b, i += 64) {
            a0 = _mm_sub_epi8(a0, _mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i)), nl));
            a1 = _mm_sub_epi8(a1, _mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i + 16)), nl));
            a2 = _mm_sub_epi8(a2, _mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i + 32)), nl));
            a3 = _mm_sub_epi8(a3, _mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)(p + i + 48)), nl));
        }
        const __m128i t = _mm_add_epi64(_mm_add_epi64(_mm_sad_epu8(a0, zero), _mm_sad_epu8(a1, zero)),
                                        _mm_add_epi64(_mm_sad_epu8(a2, zero), _mm_sad_epu8(a3, zero)));
        c += (uint64_t)_mm_cvtsi128_si64(t) + (uint64_t)_mm_cvtsi128_si64(_mm_unpackhi_epi64(t, t));
    }
#endif
    for (; i < n; ++i) c += p[i] == '\n';
    return c;
}

// canonical bytes of one record (needletail write_fasta / write_fastq, LineEnding::Unix; src/utils.rs:142-154)
void canonical_record(const uint8_t* r, size_t n, uint32_t fmt, std::vector<uint8_t>& out) {
    std::vector<std::pair<size_t, size_t>> lines;  // [begin, end) without the line end
    size_t a = 0;
    while (a < n) {
        const void* q = memchr(r + a, '\n', n - a);
        size_t e = q ? (size_t)((const uint8_t*)q - r) : n;
        size_t le = e;
        if (le > a && r[le - 1] == '\r') --le;
        lines.emplace_back(a, le);
        a = e + 1;
    }
    auto put = [&](size_t b, size_t e) { out.insert(out.end(), r + b, r + e); };
    if (lines.empty()) return;
    if (fmt == SBR_FMT_FASTQ) {
        if (lines.size() < 4) { out.insert(out.end(), r, r + n); return; }
        put(lines[0].first, lines[0].second); out.push_back('\n');
        put(lines[1].first, lines[1].second); out.push_back('\n');
        out.push_back('+'); out.push_back('\n');
        put(lines[3].first, lines[3].second); out.push_back('\n');
    } else {
        put(lines[0].first, lines[0].second); out.push_back('\n');
        for (size_t i = 1; i < lines.size(); ++i) put(lines[i].first, lines[i].second);
        out.push_back('\n');
    }
}

struct Sink {  // where the records of one bin go in this pass
    FILE* fh = nullptr;
    std::vector<uint8_t> raw;
    uint64_t written_records = 0, batch_records = 0;
};

struct Gpu;
std::vector<Gpu> take_pooled(const std::vector<uint8_t>& key);
void give_pooled(const std::vector<uint8_t>& key, const std::vector<Gpu>& gpus);

class Pass {
public:
    Pass(const std::string& path, std::vector<FILE*> bin_files, const std::vector<std::string>& barcodes, uint8_t mismatch,
         Format out_format, int level, const HostOptions& opt, RunStats* stats, const std::atomic<bool>* cancel = nullptr)
        : path_(path), opt_(opt), out_format_(out_format), level_(level), stats_(stats), cancel_(cancel) {
        const auto t_setup = Clock::now();
        n_bins_ = (uint32_t)barcodes.size() + 1;
        sinks_.resize(n_bins_);
        for (uint32_t b = 0; b < n_bins_; ++b) sinks_[b].fh = bin_files[b];
        const uint32_t k = (uint32_t)barcodes[0].size();  // bc_len from row 0 (demux.rs:34, refinement H2)
        std::vector<uint8_t> tab((size_t)barcodes.size() * k, 0);
        std::vector<uint32_t> lens(barcodes.size());
        for (size_t i = 0; i < barcodes.size(); ++i) {
            const size_t len = std::min<size_t>(barcodes[i].size(), k);  // zip() never looks past seq[..bc_len]
            std::memcpy(tab.data() + i * k, barcodes[i].data(), len);
            lens[i] = (uint32_t)len;
        }
        int ndev = sbr_device_count();
        if (ndev <= 0) throw std::runtime_error(std::string("no usable GPU: ") + sbr_last_error(nullptr) + " (there is no CPU fallback)");
        G_ = std::max(1, std::min(opt.n_gpus, ndev));
        S_ = std::max<uint32_t>(1, opt.slots_per_gpu);
        cap_ = opt.batch_bytes;
        reserve_ = std::min<uint64_t>(cap_ / 8, 16ull << 20);
        gpus_.resize(G_);
        // contexts (pinned slots, device workspaces, the match table) are kept for the next pass with the same
        // table and geometry: the reverse pass of a paired-end run reuses the forward pass's
        pool_key_.assign(tab.begin(), tab.end());
        for (uint32_t v : {k, (uint32_t)mismatch, S_, (uint32_t)(cap_ >> 20), (uint32_t)G_})
            for (int j = 0; j < 4; ++j) pool_key_.push_back((uint8_t)(v >> (8 * j)));
        for (uint32_t l : lens) pool_key_.push_back((uint8_t)l);
        std::vector<Gpu> pooled = take_pooled(pool_key_);
        for (int g = 0; g < G_; ++g) {
            if ((size_t)g < pooled.size()) { gpus_[g] = pooled[g]; continue; }
            sbr_config cfg{};
            cfg.device = g; cfg.fmt = SBR_FMT_AUTO; cfg.n_barcodes = (uint32_t)barcodes.size(); cfg.bc_len = k;
            cfg.barcodes = tab.data(); cfg.bc_lens = lens.data(); cfg.mismatch = mismatch; cfg.n_slots = S_;
            cfg.batch_bytes = cap_; cfg.max_records = 0; cfg.flags = 0;
            if (sbr_create(&gpus_[g].ctx, &cfg) != SBR_OK)
                throw std::runtime_error(std::string("sbr_create failed: ") + sbr_last_error(nullptr));
            gpus_[g].slot_host.resize(S_);
            for (uint32_t s = 0; s < S_; ++s) {
                uint64_t c = 0;
                if (sbr_slot_buffer(gpus_[g].ctx, s, &gpus_[g].slot_host[s], &c) != SBR_OK)
                    throw std::runtime_error(std::string("sbr_slot_buffer failed: ") + sbr_last_error(gpus_[g].ctx));
            }
        }
        if (stats_) stats_->setup_s += secs(t_setup, Clock::now());
    }
    ~Pass() {
        if (!error_ && !gpus_.empty() && gpus_.back().ctx) { give_pooled(pool_key_, gpus_); return; }
        for (auto& g : gpus_) sbr_destroy(g.ctx);  // after a failure nothing is assumed about the contexts' state
    }

    // strict: SE semantics (`record?`, demux.rs:50: a bad record is an error); otherwise PE semantics (the loop
    // `while let Some(Ok(record))` just stops, demux.rs:101,114).  Returns is_unk_empty.
    bool run(bool strict, Counts& nb_records, const std::vector<std::string>& barcodes) {
        const auto t0 = Clock::now();
        std::thread reader([&] { reader_main(); });
        std::thread writer([&] { writer_main(); });
        try {
            gpu_main();
        } catch (...) {
            fail(std::current_exception());
        }
        to_write_.close();
        reader.join();
        writer.join();
        if (stats_) stats_->total_s += secs(t0, Clock::now());
        if (error_) std::rethrow_exception(error_);
        for (uint32_t b = 0; b + 1 < n_bins_; ++b)
            if (sinks_[b].written_records) {
                auto it = std::find_if(nb_records.begin(), nb_records.end(), [&](auto& p) { return p.first == barcodes[b]; });
                if (it == nb_records.end()) nb_records.emplace_back(barcodes[b], sinks_[b].written_records);
                else it->second += sinks_[b].written_records;
            }
        const bool unk_empty = sinks_[n_bins_ - 1].written_records == 0;
        if (bad_flags_) {
            char msg[160];
            snprintf(msg, sizeof msg, "record %llu: %s%s%s%s%s", (unsigned long long)bad_record_,
                     (bad_flags_ & SBR_ST_INVALID_START) ? "invalid start byte " : "",
                     (bad_flags_ & SBR_ST_INVALID_SEP) ? "invalid separator line " : "",
                     (bad_flags_ & SBR_ST_UNEQUAL_LEN) ? "sequence and quality lengths differ " : "",
                     (bad_flags_ & SBR_ST_SHORT_READ) ? "read shorter than the barcode " : "",
                     (bad_flags_ & SBR_ST_TRUNCATED) ? "truncated record" : "");
            if (bad_flags_ & SBR_ST_SHORT_READ) throw ShortReadPanic(msg);
            if (strict) throw ParseError(msg);
        }
        return unk_empty;
    }

private:
    void fail(std::exception_ptr e) {
        {
            std::lock_guard<std::mutex> l(err_m_);
            if (!error_) error_ = e;
        }
        stop_.store(true);
        slot_cv_.notify_all();
        filled_.close();
        to_write_.close();
    }

    // ---- reader: decompress into the next slot, cut at a guessed boundary --------------------------------------
    void reader_main() {
        try {
            Reader in(path_);
            std::vector<uint8_t> carry;  // bytes behind the previous batch's cut
            uint64_t seq = 0;
            uint32_t fmt = 0;
            bool eof = false;
            while (!eof && !stop_.load()) {
                if (cancelled()) { stop_after_error(); break; }
                {   // slot of batch `seq` is free once batch seq - G*S has been written
                    std::unique_lock<std::mutex> l(slot_m_);
                    slot_cv_.wait(l, [&] { return stop_.load() || written_ + (uint64_t)G_ * S_ > seq; });
                    if (stop_.load()) break;
                }
                const auto t0 = Clock::now();
                Batch b;
                b.seq = seq;
                b.gpu = (int)(seq % G_);
                b.slot = (uint32_t)((seq / G_) % S_);
                uint8_t* buf = gpus_[b.gpu].slot_host[b.slot];
                if (carry.size() > reserve_) throw std::runtime_error("a single record is larger than the room reserved for it (1/8 of a batch, 16 MiB at most): raise --batch-mb");
                b.off = reserve_ - carry.size();
                std::memcpy(buf + b.off, carry.data(), carry.size());
                // A batch holds at most max_records records (the library's default: batch_bytes / 32 + 16).  A FASTQ
                // record has four newlines and a FASTA record at least two, so filling stops once the newlines that
                // many records could own are in: streams of very short records (< 32 bytes) become smaller batches
                // instead of overflowing the record capacity.
                const uint64_t max_records = cap_ / 32 + 16;
                uint64_t nl_seen = count_newlines(carry.data(), carry.size());
                const uint64_t step = std::max<uint64_t>(64 << 10, std::min<uint64_t>(reserve_ / 2, 4ull << 20));
                std::vector<uint8_t> spill;  // read, but behind the newline budget: goes to the next batch
                uint64_t fill = reserve_;
                while (fill < cap_) {
                    size_t got = in.read(buf + fill, std::min<uint64_t>(step, cap_ - fill));
                    if (got == 0) { eof = true; break; }
                    if (fmt == 0) fmt = buf[b.off] == '>' ? SBR_FMT_FASTA : SBR_FMT_FASTQ;
                    const uint64_t budget = (max_records - 64) * (fmt == SBR_FMT_FASTA ? 2 : 4);
                    const uint64_t c = count_newlines(buf + fill, got);
                    if (nl_seen + c > budget) {
                        uint64_t keep = 0, left = budget > nl_seen ? budget - nl_seen : 0;
                        while (keep < got && left) left -= buf[fill + keep++] == '\n';
                        spill.assign(buf + fill + keep, buf + fill + got);
                        fill += keep;
                        break;
                    }
                    nl_seen += c;
                    fill += got;
                }
                if (!eof && spill.empty()) {  // learn about the end of the stream before deciding whether this batch is final
                    uint8_t probe;
                    if (in.read(&probe, 1) == 0) eof = true;
                    else { peeked_ = true; peek_byte_ = probe; }
                }
                uint64_t n = fill - b.off;
                // needletail::parse_fastx_reader(..)? (demux.rs:23, 85-86) fails before any record is read when the stream
                // is empty or starts with neither '>' nor '@': an error in single-end AND paired-end mode
                if (seq == 0 && n == 0) throw ParseError("record 0: empty input");
                if (seq == 0 && buf[b.off] != '>' && buf[b.off] != '@') throw ParseError("record 0: invalid start byte");
                carry.clear();
                if (!eof) {
                    uint64_t cut = guess_cut(buf + b.off, n, fmt, std::min<uint64_t>(spill.empty() ? reserve_ : reserve_ / 2, 1ull << 20));
                    if (cut == 0) cut = n;  // no boundary in sight: let the scan decide
                    carry.assign(buf + b.off + cut, buf + b.off + n);
                    carry.insert(carry.end(), spill.begin(), spill.end());
                    if (peeked_) { carry.push_back(peek_byte_); peeked_ = false; }
                    // FASTA: a line that starts with '>' IS a record start (no guess involved), so the batch in front
                    // of it is submitted as "ends at a record boundary"; a non-final FASTA batch would otherwise hold
                    // its last record back as possibly incomplete
                    b.exact_end = fmt == SBR_FMT_FASTA && cut < n;
                    n = cut;
                }
                b.n = n;
                b.final_batch = eof;
                if (stats_) { stats_->read_s += secs(t0, Clock::now()); stats_->bytes_in += n; }
                if (n == 0 && eof && seq > 0) break;
                filled_.push(b);
                ++seq;
            }
        } catch (...) {
            fail(std::current_exception());
        }
        filled_.close();
    }

    // ---- GPU thread: submit round-robin, wait in order, verify the guessed cuts ----------------------------------
    void submit(Batch& b) {
        Gpu& g = gpus_[b.gpu];
        uint8_t* buf = g.slot_host[b.slot];
        if (!prefix_.empty()) {  // the unconsumed tail of the previous batch goes in front
            if (prefix_.size() > b.off) throw std::runtime_error("unconsumed tail does not fit in front of the next batch");
            b.off -= prefix_.size();
            b.n += prefix_.size();
            std::memcpy(buf + b.off, prefix_.data(), prefix_.size());
            prefix_.clear();
        }
        if (sbr_submit(g.ctx, b.slot, b.off, b.n, b.seq, (b.final_batch || b.exact_end) ? 1 : 0) != SBR_OK)
            throw std::runtime_error(std::string("sbr_submit failed: ") + sbr_last_error(g.ctx));
    }
    void wait(Batch& b) {
        const auto t0 = Clock::now();
        Gpu& g = gpus_[b.gpu];
        if (sbr_wait(g.ctx, b.slot, &b.res) != SBR_OK)
            throw std::runtime_error(std::string("sbr_wait failed: ") + sbr_last_error(g.ctx));
        if (stats_) stats_->gpu_wait_s += secs(t0, Clock::now());
    }
    void gpu_main() {
        std::deque<Batch> flight;
        bool input_done = false;
        while (!stop_.load()) {
            if (cancelled()) { stop_after_error(); break; }
            Batch nb;
            // keep every slot busy: take whatever the reader has ready
            while (!input_done && flight.size() < (size_t)G_ * S_ && filled_.try_pop(nb)) {
                submit(nb);
                flight.push_back(nb);
            }
            if (flight.empty()) {
                if (!filled_.pop(nb)) break;  // closed and drained
                submit(nb);
                flight.push_back(nb);
                continue;
            }
            Batch b = flight.front();
            flight.pop_front();
            wait(b);
            const bool errored = (b.res.status_flags & SBR_ST_ERROR_MASK) != 0;
            bool blank_tail = false;
            if (!errored && b.res.consumed < b.n && b.final_batch && !(b.res.status_flags & SBR_ST_REC_OVERFLOW)) {
                // end of the stream: what the library leaves behind without an error are the blank lines needletail
                // tolerates after the last record (sbr_submit drops up to three of them)
                const uint8_t* buf = gpus_[b.gpu].slot_host[b.slot];
                blank_tail = true;
                // (sbr_submit pads the text it copies with zero bytes: the first 16 bytes of the dropped tail may read 0)
                for (uint64_t q = b.off + b.res.consumed; q < b.off + b.n; ++q)
                    blank_tail = blank_tail && (buf[q] == '\n' || buf[q] == '\r' || buf[q] == 0);
                if (!blank_tail) throw std::runtime_error("internal: final batch left bytes unconsumed without an error");
            }
            if (!errored && b.res.consumed < b.n && !blank_tail) {
                // the guess was wrong (or the record capacity was hit): the rest belongs in front of the next batch
                const uint8_t* buf = gpus_[b.gpu].slot_host[b.slot];
                prefix_.assign(buf + b.off + b.res.consumed, buf + b.off + b.n);
                if (b.res.consumed == 0) throw std::runtime_error("no progress: a single record is larger than a batch");
                if (stats_) ++stats_->respeculated;
                if (!flight.empty()) {  // the next batch already ran from a wrong start: run it again
                    Batch& nx = flight.front();
                    sbr_result dump;
                    if (sbr_wait(gpus_[nx.gpu].ctx, nx.slot, &dump) != SBR_OK)
                        throw std::runtime_error(std::string("sbr_wait failed: ") + sbr_last_error(gpus_[nx.gpu].ctx));
                    submit(nx);
                } else if (b.final_batch) {  // capacity overflow on the last batch: feed the rest as one more batch
                    Batch extra = b;
                    extra.seq = b.seq + 1;  // reuses the same slot after the writer is done with it: serialise
                    pending_extra_ = true;
                    extra_ = extra;
                }
            }
            if (stats_) { ++stats_->batches; stats_->records += b.res.n_records; }
            to_write_.push(b);
            if (errored) { stop_after_error(); break; }
            if (pending_extra_) run_extra();
        }
        // drain: results of batches behind an error are not wanted, but their slots must be quiet before teardown
        for (auto& b : flight) {
            sbr_result dump;
            sbr_wait(gpus_[b.gpu].ctx, b.slot, &dump);
        }
    }
    void stop_after_error() {
        stop_.store(true);
        slot_cv_.notify_all();
    }
    bool cancelled() const { return cancel_ && cancel_->load(); }
    void run_extra() {  // rare: the last batch hit the record capacity; its rest is processed after the writer released the slot
        pending_extra_ = false;
        while (!prefix_.empty()) {
            Batch b = extra_;
            {
                std::unique_lock<std::mutex> l(slot_m_);
                slot_cv_.wait(l, [&] { return stop_.load() || written_ >= b.seq; });
                if (stop_.load()) return;
            }
            uint8_t* buf = gpus_[b.gpu].slot_host[b.slot];
            if (prefix_.size() > cap_ - reserve_) throw std::runtime_error("records too small for the configured record capacity");
            b.off = reserve_;
            b.n = prefix_.size();
            std::memcpy(buf + b.off, prefix_.data(), prefix_.size());
            prefix_.clear();
            b.final_batch = true;
            submit(b);
            wait(b);
            if (!(b.res.status_flags & SBR_ST_ERROR_MASK) && b.res.consumed < b.n && (b.res.status_flags & SBR_ST_REC_OVERFLOW)) {
                prefix_.assign(buf + b.off + b.res.consumed, buf + b.off + b.n);  // (else: the stream's blank tail)
                extra_.seq = b.seq + 1;
            }
            if (stats_) { ++stats_->batches; stats_->records += b.res.n_records; }
            to_write_.push(b);
            if (b.res.status_flags & SBR_ST_ERROR_MASK) { stop_after_error(); return; }
        }
    }

    // ---- writer: batches in order; gather per bin and compress per 4 MB member in parallel, append in order --------
    void gather_bin(const Batch& b, uint32_t bin, uint32_t nok) {
        const sbr_result& r = b.res;
        Sink& s = sinks_[bin];
        s.raw.clear();
        s.batch_records = 0;
        const uint32_t lo = r.bin_start[bin], hi = r.bin_start[bin + 1];
        if (lo == hi) return;
        const uint8_t* text = gpus_[b.gpu].slot_host[b.slot] + b.off;
        const bool check = r.rec_key != nullptr;  // some record of the batch needs re-serialisation
        s.raw.reserve((size_t)(r.rec_off[r.order[hi - 1] + 1] - r.rec_off[r.order[lo]]) / 4 + 4096);
        for (uint32_t j = lo; j < hi; ++j) {
            const uint32_t i = r.order[j];
            if (i >= nok) continue;  // at or behind the first bad record: the reference never got there
            const uint8_t* rec = text + r.rec_off[i];
            const size_t len = r.rec_off[i + 1] - r.rec_off[i];
            if (check && (r.fmt == SBR_FMT_FASTA || (r.rec_key[i] & (1ull << 49)))) canonical_record(rec, len, r.fmt, s.raw);
            else s.raw.insert(s.raw.end(), rec, rec + len);
            ++s.batch_records;
        }
    }
    template <class F>
    void parallel_for(uint32_t n, int nthreads, F&& body) {
        std::atomic<uint32_t> next{0};
        std::exception_ptr werr;
        std::mutex wm;
        auto work = [&] {
            try {
                for (uint32_t i; (i = next.fetch_add(1)) < n;) body(i);
            } catch (...) {
                std::lock_guard<std::mutex> l(wm);
                werr = std::current_exception();
            }
        };
        const int nt = (int)std::min<uint32_t>((uint32_t)nthreads, n);
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        if (werr) std::rethrow_exception(werr);
    }
    void writer_main() {
        try {
            const int nthreads = opt_.writer_threads > 0 ? opt_.writer_threads : (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
            constexpr size_t MEMBER = 4u << 20;  // raw bytes per compressed member
            struct Piece { uint32_t bin; size_t lo, hi; std::vector<uint8_t> packed; };
            Batch b;
            while (to_write_.pop(b)) {
                const auto t0 = Clock::now();
                const sbr_result& r = b.res;
                uint32_t nok = r.n_records;
                if (r.status_flags & SBR_ST_ERROR_MASK) {
                    nok = std::min(nok, r.first_bad_record);
                    bad_flags_ = r.first_bad_flags;
                    bad_record_ = records_before_ + r.first_bad_record;
                }
                // uncompressed output: every bin is its own file, so the thread that gathered a bin also appends it
                parallel_for(n_bins_, nthreads, [&](uint32_t bin) {
                    gather_bin(b, bin, nok);
                    Sink& s = sinks_[bin];
                    if (out_format_ == Format::No && !s.raw.empty() && fwrite(s.raw.data(), 1, s.raw.size(), s.fh) != s.raw.size())
                        throw std::runtime_error("write failed");
                });
                if (out_format_ != Format::No) {
                    std::vector<Piece> pieces;  // in (bin, offset) order: members of a file are appended in this order
                    for (uint32_t bin = 0; bin < n_bins_; ++bin)
                        for (size_t lo = 0; lo < sinks_[bin].raw.size(); lo += MEMBER)
                            pieces.push_back(Piece{bin, lo, std::min(lo + MEMBER, sinks_[bin].raw.size()), {}});
                    parallel_for((uint32_t)pieces.size(), nthreads, [&](uint32_t k) {
                        Piece& pc = pieces[k];
                        compress_member(out_format_, level_, sinks_[pc.bin].raw.data() + pc.lo, pc.hi - pc.lo, pc.packed);
                    });
                    for (auto& pc : pieces)
                        if (fwrite(pc.packed.data(), 1, pc.packed.size(), sinks_[pc.bin].fh) != pc.packed.size())
                            throw std::runtime_error("write failed");
                }
                for (auto& s : sinks_) s.written_records += s.batch_records;
                records_before_ += r.n_records;
                if (stats_) stats_->write_s += secs(t0, Clock::now());
                {
                    std::lock_guard<std::mutex> l(slot_m_);
                    written_ = b.seq + 1;
                }
                slot_cv_.notify_all();
            }
            for (auto& s : sinks_) fflush(s.fh);
        } catch (...) {
            fail(std::current_exception());
        }
    }

    std::vector<uint8_t> pool_key_;
    std::string path_;
    HostOptions opt_;
    Format out_format_;
    int level_;
    RunStats* stats_;
    const std::atomic<bool>* cancel_ = nullptr;  // paired-end: the other mate's pass failed, stop quietly
    uint32_t n_bins_ = 0;
    int G_ = 1;
    uint32_t S_ = 2;
    uint64_t cap_ = 0, reserve_ = 0;
    std::vector<Gpu> gpus_;
    std::vector<Sink> sinks_;
    Channel<Batch> filled_, to_write_;
    std::mutex slot_m_;
    std::condition_variable slot_cv_;
    uint64_t written_ = 0;  // batches fully written (== next sequence number the writer expects)
    std::atomic<bool> stop_{false};
    std::mutex err_m_;
    std::exception_ptr error_;
    std::vector<uint8_t> prefix_;
    bool peeked_ = false;
    uint8_t peek_byte_ = 0;
    bool pending_extra_ = false;
    Batch extra_;
    uint32_t bad_flags_ = 0;
    uint64_t bad_record_ = 0, records_before_ = 0;
};

// one parked set of contexts (a process runs one table at a time)
std::mutex g_pool_m;
std::vector<uint8_t> g_pool_key;
std::vector<Gpu> g_pool;
std::vector<Gpu> take_pooled(const std::vector<uint8_t>& key) {
    std::lock_guard<std::mutex> l(g_pool_m);
    std::vector<Gpu> out;
    if (!g_pool.empty() && g_pool_key == key) out.swap(g_pool);
    else {
        for (auto& g : g_pool) sbr_destroy(g.ctx);
        g_pool.clear();
    }
    return out;
}
void give_pooled(const std::vector<uint8_t>& key, const std::vector<Gpu>& gpus) {
    std::lock_guard<std::mutex> l(g_pool_m);
    for (auto& g : g_pool) sbr_destroy(g.ctx);
    g_pool = gpus;
    g_pool_key = key;
}

void truncate_file(FILE* fh) {
    if (!fh) return;
    fflush(fh);
    if (ftruncate(fileno(fh), 0) != 0) throw std::runtime_error("ftruncate failed");
    rewind(fh);
}

std::vector<std::string> barcode_list(const Barcode& bd) {
    if (bd.rows.empty()) throw std::runtime_error("Barcode data is empty");  // demux.rs:31-33
    std::vector<std::string> v;
    for (auto& r : bd.rows) v.push_back(r.barcode);
    return v;
}

}  // namespace

namespace detail {
uint64_t host_guess_cut(const uint8_t* p, uint64_t n, uint32_t fmt, uint64_t window) { return guess_cut(p, n, fmt, window); }
uint64_t host_count_newlines(const uint8_t* p, uint64_t n) { return count_newlines(p, n); }
void host_canonical_record(const uint8_t* rec, size_t n, uint32_t fmt, std::vector<uint8_t>& out) { canonical_record(rec, n, fmt, out); }
}  // namespace detail

std::pair<Counts&, bool> se_demux(const std::string& file, Format format, int level, Barcode& bd, uint8_t mismatch,
                                  Counts& nb_records, const HostOptions& opt, RunStats* stats) {
    const Format original = which_format(file);
    if (format == Format::No) format = original;  // demux.rs:26-29
    const auto bcs = barcode_list(bd);
    if (bd.unknown.empty()) throw std::runtime_error("Missing 'XXX' fallback barcode entry in barcode_data");  // demux.rs:43-46
    std::vector<FILE*> files;
    for (auto& r : bd.rows) files.push_back(r.files.at(0).fh);
    files.push_back(bd.unknown.at(0).fh);
    Pass pass(file, files, bcs, mismatch, format, level, opt, stats);
    const bool unk_empty = pass.run(/*strict=*/true, nb_records, bcs);
    return {nb_records, unk_empty};
}

std::pair<Counts&, std::string> pe_demux(const std::string& forward, const std::string& reverse, Format format, int level,
                                         Barcode& bd, uint8_t mismatch, Counts& nb_records, const HostOptions& opt,
                                         RunStats* stats) {
    Format compression = which_format(forward);  // both mates are written with the forward-derived format (demux.rs:81,95-97)
    (void)which_format(reverse);
    if (format != Format::No) compression = format;
    const auto bcs = barcode_list(bd);
    if (bd.unknown.size() < 2) throw std::runtime_error("Missing 'XXX' fallback barcode entry in barcode_data");
    // The reference runs the forward pass, then the reverse pass (demux.rs:101-110, 114-123).  The mates are
    // independent streams writing to disjoint files, so both passes may run at once (each with its own contexts):
    // with compressed inputs the single-threaded inflate of each file is the wall, and two of them run in parallel.
    // Both parsers are built before the first record is processed (demux.rs:85-86): an empty mate, or one that starts
    // with neither '>' nor '@', fails the whole call and nothing is written for either mate.
    for (const std::string* path : {&forward, &reverse}) {
        Reader probe(*path);
        uint8_t first = 0;
        if (probe.read(&first, 1) == 0) throw ParseError(*path + ": empty input");
        if (first != '>' && first != '@') throw ParseError(*path + ": invalid start byte");
    }
    bool unk[2] = {true, true};
    Counts counts[2];
    RunStats st[2];
    std::exception_ptr errs[2];
    std::atomic<bool> cancel_reverse{false};
    auto run_mate = [&](int mate) {
        try {
            std::vector<FILE*> files;
            for (auto& r : bd.rows) files.push_back(r.files.at(mate).fh);
            files.push_back(bd.unknown.at(mate).fh);
            Pass pass(mate == 0 ? forward : reverse, files, bcs, mismatch, compression, level, opt, stats ? &st[mate] : nullptr,
                      mate == 1 ? &cancel_reverse : nullptr);
            unk[mate] = pass.run(/*strict=*/false, counts[mate], bcs);
        } catch (...) {
            errs[mate] = std::current_exception();
            if (mate == 0) cancel_reverse.store(true);
        }
    };
    if (opt.parallel_mates) {
        std::thread t1([&] { run_mate(1); });
        run_mate(0);
        t1.join();
        if (errs[0]) {
            // The reference reaches the reverse loop only after the forward loop has finished (demux.rs:101-123): when
            // the forward pass dies (short-read panic, I/O error) the reverse files are still empty.  Here the reverse
            // pass ran beside it: it has been told to stop, and what it wrote is taken back.
            for (auto& r : bd.rows) truncate_file(r.files.at(1).fh);
            truncate_file(bd.unknown.at(1).fh);
        }
    } else {
        run_mate(0);
        if (!errs[0]) run_mate(1);
    }
    for (int mate = 0; mate < 2; ++mate)
        if (errs[mate]) std::rethrow_exception(errs[mate]);
    for (int mate = 0; mate < 2; ++mate) {  // forward counts first, then reverse: the order the sequential passes add them
        for (auto& kv : counts[mate]) {
            auto it = std::find_if(nb_records.begin(), nb_records.end(), [&](auto& p) { return p.first == kv.first; });
            if (it == nb_records.end()) nb_records.push_back(kv);
            else it->second += kv.second;
        }
        if (stats) {
            stats->read_s += st[mate].read_s; stats->gpu_wait_s += st[mate].gpu_wait_s; stats->write_s += st[mate].write_s;
            stats->setup_s = std::max(stats->setup_s, st[mate].setup_s);
            stats->total_s = std::max(stats->total_s, st[mate].total_s) + (opt.parallel_mates ? 0.0 : (mate ? st[0].total_s : 0.0));
            stats->bytes_in += st[mate].bytes_in; stats->records += st[mate].records; stats->batches += st[mate].batches;
            stats->respeculated += st[mate].respeculated;
        }
    }
    return {nb_records, std::string(unk[0] ? "true" : "false") + (unk[1] ? "true" : "false")};
}

}  // namespace sabreur
