// demux_host.hpp -- the reference's demux entry points over the C ABI (include/sabreur_gpu.h).
//
//   se_demux / pe_demux     src/demux.rs:13-67 / 71-128: same argument meaning and the same results
//                           ((nb_records, is_unk_empty) / (nb_records, "truefalse"...)); the per-record loop
//                           body (find first matching barcode, route the record) runs on the GPU(s).
//   Barcode                 src/demux.rs:10: barcode -> output files.  Here an ordered table: candidate order is
//                           the barcode file's row order (refinement H1), the "XXX" entry is `unknown`.
//
// Nothing in this host matches reads on the CPU: without libsabreur_gpu.so and a GPU every call fails.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "codec.hpp"

namespace sabreur {

struct OutFile {
    std::string path;
    FILE* fh = nullptr;  // opened create+append like OpenOptions::new().create(true).append(true) (src/main.rs:87-90)
};

struct BarcodeRow {
    std::string barcode;         // raw bytes of column 1
    std::vector<OutFile> files;  // [R1] or [R1, R2]
};

struct Barcode {
    std::vector<BarcodeRow> rows;  // barcode-file order, duplicates already folded (the last row's files win, main.rs:100,149)
    std::vector<OutFile> unknown;  // the "XXX" entry (main.rs:152-158 / 103-116)
};

// ordered stand-in of HashMap<&[u8], u32>: only barcodes with >= 1 hit appear (entry().or_insert, demux.rs:58)
using Counts = std::vector<std::pair<std::string, uint64_t>>;

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct ShortReadPanic : std::runtime_error {  // the slice panic of src/demux.rs:54,102,115
    using std::runtime_error::runtime_error;
};

struct HostOptions {
    int n_gpus = 1;                 // batches are dealt round-robin to this many GPUs
    uint64_t batch_bytes = 256ull << 20;
    uint32_t slots_per_gpu = 2;
    int writer_threads = 0;         // 0 = hardware concurrency (capped)
    bool parallel_mates = true;     // paired-end: run the forward and the reverse pass at the same time
    bool verbose = false;
};

struct RunStats {  // where the wall time went (reported by the CLI with --timing)
    double read_s = 0, gpu_wait_s = 0, write_s = 0, total_s = 0;
    double setup_s = 0;  // contexts, pinned slots, device workspaces, the match table (before the first byte is read)
    uint64_t bytes_in = 0, records = 0, batches = 0, respeculated = 0;
};

std::pair<Counts&, bool> se_demux(const std::string& file, Format format, int level, Barcode& barcode_data, uint8_t mismatch,
                                  Counts& nb_records, const HostOptions& opt = HostOptions(), RunStats* stats = nullptr);

std::pair<Counts&, std::string> pe_demux(const std::string& forward, const std::string& reverse, Format format, int level,
                                         Barcode& barcode_data, uint8_t mismatch, Counts& nb_records,
                                         const HostOptions& opt = HostOptions(), RunStats* stats = nullptr);

// src/utils.rs:103-112
std::vector<std::vector<std::string>> split_by_tab(const std::string& content);

// pure pieces of the host pipeline, exposed for the CPU self-test (host_selftest.cpp, tests/test_host_cpu.py)
namespace detail {
// offset where the last (possibly incomplete) record of [p, p + n) is guessed to start, looking back `window` bytes
uint64_t host_guess_cut(const uint8_t* p, uint64_t n, uint32_t fmt, uint64_t window);
uint64_t host_count_newlines(const uint8_t* p, uint64_t n);
// needletail's canonical bytes of one record (src/utils.rs:142-154), appended to `out`
void host_canonical_record(const uint8_t* rec, size_t n, uint32_t fmt, std::vector<uint8_t>& out);
}  // namespace detail

}  // namespace sabreur
