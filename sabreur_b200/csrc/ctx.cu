// ctx.cu -- host side of the C ABI (include/sabreur_gpu.h): contexts, slots, streams, copies.
// No CPU compute path lives here: if CUDA is unavailable every entry point fails loudly.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "scan_common.cuh"

namespace sbr {
// k_scan.cu (look-back scan: exact fallback)
size_t scan_smem_bytes();
cudaError_t scan_prepare(int device, int* grid_out);
cudaError_t scan_launch(uint32_t fmt, const uint8_t* d_text, uint32_t nbytes, uint32_t max_records, uint32_t bc_len,
                        uint32_t packed, uint32_t final_batch, uint32_t gated, BatchHeader* hdr, uint4* desc,
                        uint32_t* rec_off, uint64_t* key, int grid_cap, cudaStream_t stream);
uint32_t scan_tiles(uint32_t nbytes);
// k_scan_fast.cu (region-parallel scan)
cudaError_t scan_fast_prepare(int device, int* grid_fastq, int* grid_fasta);
size_t scan_fast_smem_bytes(uint32_t fmt);
uint32_t scan_fast_tile_bytes();
uint32_t scan_fast_regions(uint32_t nbytes, int max_regions, uint32_t* tiles_per_region, uint32_t* extra_regions);
cudaError_t scan_fast_launch(uint32_t fmt, const uint8_t* d_text, uint32_t nbytes, uint32_t max_records, uint32_t cap_r,
                             uint32_t n_regions, uint32_t tiles_per_region, uint32_t extra_regions, uint32_t bc_len,
                             uint32_t packed, uint32_t final_batch, BatchHeader* hdr, RegionInfo* info, uint32_t* region_base,
                             uint32_t* seg_rec_off, uint64_t* seg_key, uint32_t* rec_off_dense, uint4* lookback_desc,
                             uint32_t n_desc, uint32_t* stats, cudaStream_t stream);
// k_match.cu
struct MatchLaunch {
    const uint64_t* key_dense; const uint64_t* seg_key; const uint32_t* seg_rec_off; const uint32_t* region_base;
    uint32_t cap_r;
    const BatchHeader* hdr;
    const uint2* tab; const uint16_t* lut; const uint16_t* lut1; uint32_t uniform; const uint8_t* rows; const uint32_t* lens;
    uint32_t n, m, k, stride;
    const uint8_t* text; uint32_t nbytes;
    uint16_t* assign; uint32_t* rec_off; uint64_t* key_out; uint32_t* cta_cnt; uint32_t nbins;
    int grid;
    int fused; uint32_t nbits; uint32_t* bin_count; uint32_t* bin_start; uint32_t* order; BatchHeader* zero_next;
    uint32_t force_keys;
    uint32_t lean;
};
size_t match_smem(uint32_t n, uint32_t stride, uint32_t nbins, bool generic, bool lean);
cudaError_t match_prepare(bool max_shared);
int match_fused_max_grid(int device, int mode, size_t smem, bool lean);
cudaError_t match_build_lut(const uint2* d_tab, uint32_t n, uint32_t m, uint32_t k, uint16_t* d_lut, cudaStream_t stream);
cudaError_t match_launch(const MatchLaunch& L, cudaStream_t stream);
// k_partition.cu
struct PartitionPlan {
    int grid;
    uint32_t V;
    size_t smem;
    uint32_t nbits;
};
PartitionPlan partition_plan(uint32_t nbins, uint32_t max_records, int sms);
cudaError_t partition_prepare(bool max_shared);
cudaError_t partition_launch(const uint16_t* d_assign, const BatchHeader* d_hdr, uint32_t nbins,
                             uint32_t* d_cta_cnt, uint32_t* d_bin_count, uint32_t* d_bin_start, uint32_t* d_order,
                             uint32_t* d_done_counter, BatchHeader* d_zero_next, const PartitionPlan& p, cudaStream_t stream);
}  // namespace sbr

using namespace sbr;

static thread_local std::string g_create_error;

namespace {

constexpr int TIMING_RING = 2048;
constexpr int EV_PER = 5;

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t hdr_ready = nullptr;
    // overlapped mode (SBR_CFG_OVERLAP, device-resident batches): match + partition run on `aux` beside the NEXT batch's scan
    cudaStream_t aux = nullptr;
    cudaEvent_t scan_done = nullptr;  // recorded behind the scan on the caller's stream; `aux` waits for it
    cudaEvent_t mp_done = nullptr;    // recorded behind match + partition on `aux`; the slot's next scan waits for it
    bool mp_pending = false;
    // device
    uint8_t* d_text = nullptr;
    BatchHeader* d_hdr = nullptr;   // the header of the batch in flight: one of d_hdr2[0..1], alternating
    BatchHeader* d_hdr2 = nullptr;  // both start out zeroed; the last kernel of a batch clears the other one for the next batch
    uint32_t hdr_idx = 0;
    uint4* d_desc = nullptr;
    uint32_t* d_rec_off = nullptr;
    uint64_t* d_key = nullptr;
    uint16_t* d_assign = nullptr;
    uint32_t* d_cta_cnt = nullptr;
    uint32_t* d_bin_count = nullptr;
    uint32_t* d_bin_start = nullptr;
    uint32_t* d_order = nullptr;
    RegionInfo* d_info = nullptr;
    uint32_t* d_region_base = nullptr;
    uint32_t* d_seg_rec_off = nullptr;
    uint64_t* d_seg_key = nullptr;
    uint32_t* d_done = nullptr;  // k_binscan's last-block counter (self-resetting)
    // pinned host
    uint8_t* h_text = nullptr;
    BatchHeader* h_hdr = nullptr;
    uint32_t* h_rec_off = nullptr;
    uint64_t* h_key = nullptr;
    uint16_t* h_assign = nullptr;
    uint32_t* h_bin_count = nullptr;
    uint32_t* h_bin_start = nullptr;
    uint32_t* h_order = nullptr;
    // state of the batch in flight
    bool in_flight = false;
    bool host_error = false;  // batch rejected before launch (bad first byte)
    uint32_t host_error_bits = 0;
    uint64_t batch_seq = 0;
    uint64_t nbytes_in = 0;   // as submitted
    uint32_t nbytes_eff = 0;  // after final-batch tail normalisation
    uint32_t fmt = 0;
    int final_batch = 0;
    cudaStream_t last_stream = nullptr;
    bool via_submit = false;  // results' header copy already queued behind hdr_ready
    bool supplied_nl = false; // the final newline of the stream was missing and has been supplied
    const uint8_t* cur_text = nullptr;  // device text of the batch in flight (for a re-run through the exact scan)
    uint32_t cur_cap_r = 0;
};

}  // namespace

struct sbr_ctx {
    sbr_config cfg{};
    int device = 0;
    int sms = 0;
    int scan_grid = 0;
    int fast_grid_q = 0, fast_grid_a = 0;  // CTAs (= regions at most) of the region-parallel scan, FASTQ / FASTA
    bool exact_only = false;
    uint32_t nbins = 0;
    uint32_t max_records = 0;
    uint32_t max_tiles = 0;
    bool packed = false;
    std::vector<uint8_t> rows;     // n * stride raw bytes
    std::vector<uint32_t> lens;    // significant bytes per row
    uint2* d_tab = nullptr;        // packed rows {codes, care}
    uint8_t* d_rows = nullptr;     // raw rows (generic path)
    uint32_t* d_lens = nullptr;
    uint16_t* d_lut = nullptr;
    uint16_t* d_lut1 = nullptr;  // first-match table for mismatch - 1 (reads with one non-ACGT base)
    bool uniform = false;        // every row spans the whole match window
    uint64_t lut_entries = 0;
    PartitionPlan plan{};
    bool fused = false;  // match + partition in one cooperative launch
    bool overlap = false;  // device-resident batches: match + partition on the slot's aux stream, beside the next scan
    bool lean = false;     // the match kernel's shared-memory diet (16-bit per-warp counters): big tables beside the scan
    uint32_t* d_stats = nullptr;  // {batches the region-parallel scan finished, batches it handed to the exact scan, 0, 0}
    std::vector<Slot> slots;
    std::string err;
    // timing ring
    bool timing = false;
    std::vector<cudaEvent_t> ev;  // EV_PER per batch: scan start, scan end, match end, partition end, match start (overlapped mode)
    int ev_used = 0;
    uint64_t launches = 0;
};

static int fail(sbr_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(c, call)                                                                                       \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail((c), e__ == cudaErrorMemoryAllocation ? SBR_E_NOMEM : SBR_E_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e__));                                                         \
    } while (0)

extern "C" int sbr_abi_version(void) { return SBR_ABI_VERSION; }

extern "C" int sbr_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        fail(nullptr, SBR_E_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return SBR_E_NO_DEVICE;
    }
    return n;
}

extern "C" const char* sbr_last_error(const sbr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

static void free_slot(Slot& s) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.d_text); cudaFree(s.d_hdr2); cudaFree(s.d_desc); cudaFree(s.d_rec_off); cudaFree(s.d_key);
    cudaFree(s.d_done); cudaFree(s.d_info); cudaFree(s.d_region_base); cudaFree(s.d_seg_rec_off); cudaFree(s.d_seg_key);
    cudaFree(s.d_assign); cudaFree(s.d_cta_cnt); cudaFree(s.d_bin_count); cudaFree(s.d_bin_start); cudaFree(s.d_order);
    cudaFreeHost(s.h_text); cudaFreeHost(s.h_hdr); cudaFreeHost(s.h_rec_off); cudaFreeHost(s.h_key);
    cudaFreeHost(s.h_assign); cudaFreeHost(s.h_bin_count); cudaFreeHost(s.h_bin_start); cudaFreeHost(s.h_order);
    if (s.aux) cudaStreamSynchronize(s.aux);
    if (s.hdr_ready) cudaEventDestroy(s.hdr_ready);
    if (s.scan_done) cudaEventDestroy(s.scan_done);
    if (s.mp_done) cudaEventDestroy(s.mp_done);
    if (s.aux) cudaStreamDestroy(s.aux);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
}

extern "C" void sbr_destroy(sbr_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& s : c->slots) free_slot(s);
    for (auto e : c->ev) cudaEventDestroy(e);
    cudaFree(c->d_tab); cudaFree(c->d_rows); cudaFree(c->d_lens); cudaFree(c->d_lut); cudaFree(c->d_lut1);
    cudaFree(c->d_stats);
    delete c;
}

static bool is_acgt(uint8_t b) { return b == 'A' || b == 'C' || b == 'G' || b == 'T'; }

static int create_impl(sbr_ctx* c, const sbr_config* cfg) {
    c->cfg = *cfg;
    c->device = cfg->device;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(c, SBR_E_NO_DEVICE, "no CUDA device (%s): this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(c, SBR_E_INVALID_ARG, "device %d out of range", cfg->device);
    CU(c, cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU(c, cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10)
        return fail(c, SBR_E_UNSUPPORTED, "device %d is sm_%d%d; the kernels are built for sm_100a only", cfg->device,
                    prop.major, prop.minor);
    c->sms = prop.multiProcessorCount;
    const uint32_t n = cfg->n_barcodes, k = cfg->bc_len;
    if (n == 0 || k == 0 || !cfg->barcodes) return fail(c, SBR_E_INVALID_ARG, "empty barcode table");
    if (n > 6000) return fail(c, SBR_E_UNSUPPORTED, "n_barcodes %u > 6000", n);
    if (k > 64) return fail(c, SBR_E_UNSUPPORTED, "bc_len %u > 64", k);
    if (cfg->batch_bytes == 0 || cfg->batch_bytes > SCAN_MAX_BATCH)
        return fail(c, SBR_E_INVALID_ARG, "batch_bytes must be in (0, 1 GiB]");
    if (cfg->n_slots == 0 || cfg->n_slots > 64) return fail(c, SBR_E_INVALID_ARG, "n_slots must be in [1, 64]");
    if (cfg->fmt > SBR_FMT_FASTQ) return fail(c, SBR_E_INVALID_ARG, "bad fmt");
    if (cfg->mismatch > 255) return fail(c, SBR_E_INVALID_ARG, "mismatch is a u8 in the reference (src/cli.rs:33)");
    c->nbins = n + 1;
    c->rows.assign(cfg->barcodes, cfg->barcodes + (size_t)n * k);
    c->lens.resize(n);
    bool acgt = true;
    for (uint32_t b = 0; b < n; ++b) {
        uint32_t len = cfg->bc_lens ? cfg->bc_lens[b] : k;
        if (len > k) return fail(c, SBR_E_INVALID_ARG, "bc_lens[%u] = %u exceeds bc_len %u", b, len, k);
        if (b == 0 && len != k) return fail(c, SBR_E_INVALID_ARG, "row 0 defines bc_len (src/demux.rs:31-34)");
        c->lens[b] = len;
        for (uint32_t j = 0; j < len; ++j) acgt &= is_acgt(c->rows[(size_t)b * k + j]);
    }
    c->packed = acgt && k <= 16 && !(cfg->flags & SBR_CFG_GENERIC);
    c->uniform = true;
    for (uint32_t b = 0; b < n; ++b) c->uniform = c->uniform && c->lens[b] == k;
    c->max_records = cfg->max_records ? cfg->max_records : (uint32_t)(cfg->batch_bytes / 32 + 16);
    c->max_tiles = scan_tiles((uint32_t)cfg->batch_bytes) + 1;
    c->timing = (cfg->flags & SBR_CFG_TIMING) != 0;

    CU(c, scan_prepare(cfg->device, &c->scan_grid));
    CU(c, scan_fast_prepare(cfg->device, &c->fast_grid_q, &c->fast_grid_a));
    c->exact_only = (cfg->flags & SBR_CFG_EXACT_SCAN) != 0;
    c->plan = partition_plan(c->nbins, c->max_records, c->sms);
    if (match_smem(n, k, c->nbins, !c->packed, false) > 200 * 1024) return fail(c, SBR_E_UNSUPPORTED, "barcode table too large");
    const bool will_lut = c->packed && !(cfg->flags & SBR_CFG_NO_LUT) && k <= 12;
    c->overlap = (cfg->flags & SBR_CFG_OVERLAP) != 0;
    if (c->overlap) {
        // does one match CTA fit on an SM beside four CTAs of the scan?  (228 KB of shared memory per SM, 1 KB reserved per
        // CTA.)  Tables of more than ~800 bins need the lean kernel for that (16-bit counters, see k_match.cu).  If nothing
        // fits the kernels would take turns instead of sharing the SMs, which is slower than running them back to back:
        // fall back to one stream.
        size_t scan_smem = 0;
        for (uint32_t f = SBR_FMT_FASTA; f <= SBR_FMT_FASTQ; ++f)
            if (cfg->fmt == SBR_FMT_AUTO || cfg->fmt == f) scan_smem = std::max(scan_smem, scan_fast_smem_bytes(f));
        // separate launches (SBR_CFG_NO_FUSE): k_scatter keeps 32-bit per-warp cursors of its own
        const size_t scatter = (cfg->flags & SBR_CFG_NO_FUSE) ? (size_t)8 * c->nbins * 4u : 0u;
        auto fits = [&](bool lean) {
            return 4 * (scan_smem + 1024) + std::max(match_smem(n, k, c->nbins, !c->packed, lean), scatter) + 1024 <= 228 * 1024;
        };
        // 16-bit counters and offsets: the records of a whole CTA (8 warp slices, each rounded up to 32; one CTA per SM)
        // must number fewer than 65536 -- a single bin may hold all of them
        const bool lean_ok = c->packed && (uint64_t)c->max_records / (uint64_t)c->sms + 256u < 65536u;
        if (getenv("SBR_FORCE_LEAN") && lean_ok) c->lean = true;          // (tests: the lean kernel on small tables)
        else if (!fits(false) && lean_ok && fits(true)) c->lean = true;
        if (!fits(c->lean) && !getenv("SBR_OVERLAP_FORCE")) { c->overlap = false; c->lean = false; }
    }
    // (function attributes are process-wide: each context sets the split it needs, the last one created wins -- contexts of
    // both kinds in one process still run correctly, the loser just has the slower split)
    CU(c, match_prepare(c->overlap));
    CU(c, partition_prepare(c->overlap));
    if (c->overlap) {
        // match + partition of batch i run beside the scan of batch i+1: the scan holds 4 CTAs x 256 threads x 48
        // registers and ~47 KB of shared memory each on every SM, which leaves room for exactly one 256-thread CTA of
        // the match kernel (<= 64 registers, <= 34 KB).  One CTA per SM is slow when it runs alone, but it only has to
        // finish within the next scan's time.
        int per_sm = 1;
        if (const char* e = getenv("SBR_MP_PER_SM")) per_sm = atoi(e) > 0 ? atoi(e) : 1;
        const uint32_t want = (c->max_records + 256u * 4u - 1u) / (256u * 4u);
        c->plan.grid = (int)std::max(1u, std::min(want, (uint32_t)(c->sms * per_sm)));
        c->plan.V = (uint32_t)c->plan.grid * 8u;
    }
    if (!(cfg->flags & SBR_CFG_NO_FUSE)) {
        // kernel 3 rides in the match kernel's launch when the whole grid can be resident at once
        const int cap = match_fused_max_grid(cfg->device, !c->packed ? 2 : (will_lut ? 0 : 1), match_smem(n, k, c->nbins, !c->packed, c->lean), c->lean);
        if (cap >= c->sms) {
            c->fused = true;
            if (c->plan.grid > cap) { c->plan.grid = cap; c->plan.V = (uint32_t)cap * 8u; }
        }
    }

    if (c->packed) {
        std::vector<uint2> tab(n);
        for (uint32_t b = 0; b < n; ++b) {
            uint32_t codes = 0, care = 0;
            for (uint32_t j = 0; j < c->lens[b]; ++j) {
                codes |= (uint32_t)((c->rows[(size_t)b * k + j] >> 1) & 3u) << (2 * j);
                care |= 1u << (2 * j);
            }
            tab[b] = make_uint2(codes, care);
        }
        CU(c, cudaMalloc(&c->d_tab, n * sizeof(uint2)));
        CU(c, cudaMemcpy(c->d_tab, tab.data(), n * sizeof(uint2), cudaMemcpyHostToDevice));
        if (!(cfg->flags & SBR_CFG_NO_LUT) && k <= 12) {
            c->lut_entries = 1ull << (2 * k);
            CU(c, cudaMalloc(&c->d_lut, c->lut_entries * sizeof(uint16_t)));
            CU(c, match_build_lut(c->d_tab, n, cfg->mismatch, k, c->d_lut, 0));
            if (c->uniform && cfg->mismatch >= 1) {
                CU(c, cudaMalloc(&c->d_lut1, c->lut_entries * sizeof(uint16_t)));
                CU(c, match_build_lut(c->d_tab, n, cfg->mismatch - 1, k, c->d_lut1, 0));
            }
            CU(c, cudaDeviceSynchronize());
        }
    } else {
        CU(c, cudaMalloc(&c->d_rows, (size_t)n * k));
        CU(c, cudaMemcpy(c->d_rows, c->rows.data(), (size_t)n * k, cudaMemcpyHostToDevice));
        CU(c, cudaMalloc(&c->d_lens, n * sizeof(uint32_t)));
        CU(c, cudaMemcpy(c->d_lens, c->lens.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }

    CU(c, cudaMalloc(&c->d_stats, 4 * sizeof(uint32_t)));
    CU(c, cudaMemset(c->d_stats, 0, 4 * sizeof(uint32_t)));
    c->slots.resize(cfg->n_slots);
    const size_t mr = c->max_records;
    int prio_lo = 0, prio_hi = 0;
    CU(c, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (auto& s : c->slots) {
        CU(c, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU(c, cudaEventCreateWithFlags(&s.hdr_ready, cudaEventDisableTiming));
        if (c->overlap) {  // the small kernels beside the scan go first whenever they are ready
            CU(c, cudaStreamCreateWithPriority(&s.aux, cudaStreamNonBlocking, prio_hi));
            CU(c, cudaEventCreateWithFlags(&s.scan_done, cudaEventDisableTiming));
            CU(c, cudaEventCreateWithFlags(&s.mp_done, cudaEventDisableTiming));
        }
        CU(c, cudaMalloc(&s.d_hdr2, 2 * sizeof(BatchHeader)));
        CU(c, cudaMemset(s.d_hdr2, 0, 2 * sizeof(BatchHeader)));
        s.d_hdr = s.d_hdr2;
        CU(c, cudaMalloc(&s.d_desc, (size_t)c->max_tiles * sizeof(uint4)));
        CU(c, cudaMalloc(&s.d_rec_off, (mr + 1) * sizeof(uint32_t)));
        CU(c, cudaMalloc(&s.d_key, (mr + 4) * sizeof(uint64_t)));
        CU(c, cudaMalloc(&s.d_assign, (mr + 4) * sizeof(uint16_t)));
        CU(c, cudaMalloc(&s.d_cta_cnt, (size_t)c->nbins * c->plan.grid * sizeof(uint32_t)));
        CU(c, cudaMalloc(&s.d_bin_count, c->nbins * sizeof(uint32_t)));
        CU(c, cudaMalloc(&s.d_bin_start, (c->nbins + 1) * sizeof(uint32_t)));
        CU(c, cudaMalloc(&s.d_order, (mr + 1) * sizeof(uint32_t)));
        CU(c, cudaMalloc(&s.d_done, sizeof(uint32_t)));
        CU(c, cudaMemset(s.d_done, 0, sizeof(uint32_t)));
        CU(c, cudaMalloc(&s.d_info, MAX_REGIONS * sizeof(RegionInfo)));
        CU(c, cudaMalloc(&s.d_region_base, (MAX_REGIONS + 1) * sizeof(uint32_t)));
        const size_t seg = mr * 5 / 4 + 65 * (size_t)MAX_REGIONS;
        CU(c, cudaMalloc(&s.d_seg_rec_off, seg * sizeof(uint32_t)));
        CU(c, cudaMalloc(&s.d_seg_key, seg * sizeof(uint64_t)));
        CU(c, cudaMallocHost(&s.h_hdr, sizeof(BatchHeader)));
        CU(c, cudaMallocHost(&s.h_bin_count, c->nbins * sizeof(uint32_t)));
        CU(c, cudaMallocHost(&s.h_bin_start, (c->nbins + 1) * sizeof(uint32_t)));
        memset(s.h_hdr, 0, sizeof(BatchHeader));
    }
    if (c->timing) {
        c->ev.resize((size_t)TIMING_RING * EV_PER);
        for (auto& ev : c->ev) CU(c, cudaEventCreate(&ev));
    }
    return SBR_OK;
}

extern "C" int sbr_create(sbr_ctx** out, const sbr_config* cfg) {
    if (!out || !cfg) return fail(nullptr, SBR_E_INVALID_ARG, "null argument");
    *out = nullptr;
    sbr_ctx* c = new (std::nothrow) sbr_ctx();
    if (!c) return fail(nullptr, SBR_E_NOMEM, "out of host memory");
    int rc = create_impl(c, cfg);
    if (rc != SBR_OK) {
        g_create_error = c->err;
        sbr_destroy(c);
        return rc;
    }
    *out = c;
    return SBR_OK;
}

// pinned result mirrors + the slot's own text buffers are allocated on first use
static int ensure_host_mirrors(sbr_ctx* c, Slot& s) {
    if (s.h_rec_off) return SBR_OK;
    const size_t mr = c->max_records;
    CU(c, cudaMallocHost(&s.h_rec_off, (mr + 1) * sizeof(uint32_t)));
    CU(c, cudaMallocHost(&s.h_assign, (mr + 4) * sizeof(uint16_t)));
    CU(c, cudaMallocHost(&s.h_order, (mr + 1) * sizeof(uint32_t)));
    return SBR_OK;
}
static int ensure_text(sbr_ctx* c, Slot& s) {
    if (s.h_text) return SBR_OK;
    CU(c, cudaMallocHost(&s.h_text, c->cfg.batch_bytes + 64));
    CU(c, cudaMalloc(&s.d_text, c->cfg.batch_bytes + 64));
    return SBR_OK;
}

extern "C" int sbr_slot_buffer(sbr_ctx* c, uint32_t slot, uint8_t** host_ptr, uint64_t* capacity) {
    if (!c || slot >= c->slots.size() || !host_ptr) return fail(c, SBR_E_INVALID_ARG, "bad slot");
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_text(c, c->slots[slot]);
    if (rc) return rc;
    *host_ptr = c->slots[slot].h_text;
    if (capacity) *capacity = c->cfg.batch_bytes;
    return SBR_OK;
}

// scan -> match -> partition on `st`
// match (+ partition) over whatever scan result the header describes, on `st`
static int launch_match_partition(sbr_ctx* c, Slot& s, const uint8_t* d_text, uint32_t nbytes, uint32_t cap_r, cudaStream_t st,
                                  cudaEvent_t* ev) {
    MatchLaunch L{};
    L.key_dense = s.d_key; L.seg_key = s.d_seg_key; L.seg_rec_off = s.d_seg_rec_off; L.region_base = s.d_region_base;
    L.cap_r = cap_r;
    L.hdr = s.d_hdr;
    L.tab = c->packed ? c->d_tab : nullptr; L.lut = c->d_lut; L.lut1 = c->d_lut1; L.uniform = c->uniform ? 1u : 0u;
    L.rows = c->d_rows; L.lens = c->d_lens;
    L.n = c->cfg.n_barcodes; L.m = c->cfg.mismatch; L.k = c->cfg.bc_len; L.stride = c->cfg.bc_len;
    L.text = d_text; L.nbytes = nbytes;
    L.assign = s.d_assign; L.rec_off = s.d_rec_off; L.key_out = s.d_key; L.cta_cnt = s.d_cta_cnt; L.nbins = c->nbins;
    L.grid = c->plan.grid;
    L.fused = c->fused ? 1 : 0; L.nbits = c->plan.nbits;
    L.bin_count = s.d_bin_count; L.bin_start = s.d_bin_start; L.order = s.d_order;
    L.zero_next = s.d_hdr2 + (s.hdr_idx ^ 1u);
    L.force_keys = s.supplied_nl ? 1u : 0u;  // sbr_wait flags the patched last record: every scan word must be there
    L.lean = c->lean ? 1u : 0u;
    CU(c, match_launch(L, st));
    if (ev) CU(c, cudaEventRecord(ev[2], st));
    if (!c->fused)
        CU(c, partition_launch(s.d_assign, s.d_hdr, c->nbins, s.d_cta_cnt, s.d_bin_count, s.d_bin_start, s.d_order,
                               s.d_done, s.d_hdr2 + (s.hdr_idx ^ 1u), c->plan, st));
    if (ev) CU(c, cudaEventRecord(ev[3], st));
    c->launches += 1 + (c->fused ? 0 : 2);
    return SBR_OK;
}

// scan -> match -> partition on `st`.  The region-parallel scan is launched optimistically: when it hands a batch
// over (hdr->spec_fail: malformed, newline-dense, undecidable ...) the header says "no records", match and partition
// run over nothing, and sbr_wait() re-runs the batch through the exact look-back scan (rerun_exact below).  A clean
// batch therefore never pays for a launch of the fallback.
static int launch_pipeline(sbr_ctx* c, Slot& s, const uint8_t* d_text, uint32_t nbytes, uint32_t fmt, int final_batch,
                           cudaStream_t st, bool overlap) {
    cudaEvent_t* ev = nullptr;
    if (c->timing && c->ev_used < TIMING_RING) ev = &c->ev[(size_t)c->ev_used++ * EV_PER];
    // overlapped: the slot's previous match + partition (on its aux stream) still read the arrays this scan rewrites
    if (s.mp_pending) {
        if (st != s.aux) CU(c, cudaStreamWaitEvent(st, s.mp_done, 0));
        s.mp_pending = false;
    }
    // the header: the slot's two alternate; this batch takes the one the previous batch's last kernel cleared
    s.hdr_idx ^= 1u;
    s.d_hdr = s.d_hdr2 + s.hdr_idx;
    s.cur_text = d_text;
    s.cur_cap_r = 0;
    if (ev) CU(c, cudaEventRecord(ev[0], st));
    if (c->exact_only) {
        CU(c, cudaMemsetAsync(s.d_desc, 0, (size_t)scan_tiles(nbytes) * sizeof(uint4), st));
        CU(c, scan_launch(fmt, d_text, nbytes, c->max_records, c->cfg.bc_len, c->packed ? 1u : 0u, (uint32_t)final_batch, 0u,
                          s.d_hdr, s.d_desc, s.d_rec_off, s.d_key, c->scan_grid, st));
        c->launches += (fmt == SBR_FMT_FASTA ? 2 : 1);
    } else {
        uint32_t tpr = 0, extra = 0;
        uint32_t nreg = scan_fast_regions(nbytes, fmt == SBR_FMT_FASTA ? c->fast_grid_a : c->fast_grid_q, &tpr, &extra);
        // segments get 25 % + 64 records of slack over an even split: regions hold unequal record counts
        s.cur_cap_r = (uint32_t)(((uint64_t)c->max_records * 5 / 4) / nreg) + 64;
        // (the scan's epilogue clears the look-back scan's tile descriptors when it hands the batch over)
        CU(c, scan_fast_launch(fmt, d_text, nbytes, c->max_records, s.cur_cap_r, nreg, tpr, extra, c->cfg.bc_len,
                               c->packed ? 1u : 0u, (uint32_t)final_batch, s.d_hdr, s.d_info, s.d_region_base, s.d_seg_rec_off,
                               s.d_seg_key, s.d_rec_off, s.d_desc, scan_tiles(nbytes), c->d_stats, st));
        c->launches += 1;
    }
    if (ev) CU(c, cudaEventRecord(ev[1], st));
    if (ev && !overlap) CU(c, cudaEventRecord(ev[4], st));
    if (!overlap) return launch_match_partition(c, s, d_text, nbytes, s.cur_cap_r, st, ev);
    // match + partition on the slot's aux stream: the caller's stream is free for the next batch's scan at once
    CU(c, cudaEventRecord(s.scan_done, st));
    CU(c, cudaStreamWaitEvent(s.aux, s.scan_done, 0));
    if (ev) CU(c, cudaEventRecord(ev[4], s.aux));  // match + partition are timed on their own stream: ev[4] -> ev[2] -> ev[3]
    int rc = launch_match_partition(c, s, d_text, nbytes, s.cur_cap_r, s.aux, ev);
    if (rc) return rc;
    CU(c, cudaEventRecord(s.mp_done, s.aux));
    s.mp_pending = true;
    return SBR_OK;
}

// the region-parallel scan handed the batch over: exact look-back scan, then match and partition again
static int rerun_exact(sbr_ctx* c, Slot& s, cudaStream_t st) {
    CU(c, scan_launch(s.fmt, s.cur_text, s.nbytes_eff, c->max_records, c->cfg.bc_len, c->packed ? 1u : 0u,
                      (uint32_t)s.final_batch, 1u, s.d_hdr, s.d_desc, s.d_rec_off, s.d_key, c->scan_grid, st));
    c->launches += (s.fmt == SBR_FMT_FASTA ? 2 : 1);
    return launch_match_partition(c, s, s.cur_text, s.nbytes_eff, s.cur_cap_r, st, nullptr);
}

static uint32_t sniff_fmt(const sbr_ctx* c, uint8_t first) {
    if (c->cfg.fmt != SBR_FMT_AUTO) return c->cfg.fmt;
    if (first == '>') return SBR_FMT_FASTA;
    if (first == '@') return SBR_FMT_FASTQ;
    return 0;
}

extern "C" int sbr_submit(sbr_ctx* c, uint32_t slot, uint64_t offset, uint64_t nbytes, uint64_t batch_seq,
                          int final_batch) {
    if (!c || slot >= c->slots.size()) return fail(c, SBR_E_INVALID_ARG, "bad slot");
    Slot& s = c->slots[slot];
    if (s.in_flight) return fail(c, SBR_E_BUSY, "slot %u has a batch in flight", slot);
    if (offset + nbytes > c->cfg.batch_bytes) return fail(c, SBR_E_INVALID_ARG, "offset + nbytes exceeds batch_bytes");
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_text(c, s);
    if (rc) return rc;
    rc = ensure_host_mirrors(c, s);
    if (rc) return rc;
    s.batch_seq = batch_seq;
    s.nbytes_in = nbytes;
    s.final_batch = final_batch;
    s.host_error = false;
    s.host_error_bits = 0;
    s.last_stream = s.stream;
    s.via_submit = true;
    s.supplied_nl = false;
    uint64_t eff = nbytes;
    uint8_t* const h = s.h_text + offset;
    uint32_t fmt = nbytes ? sniff_fmt(c, h[0]) : (c->cfg.fmt ? c->cfg.fmt : SBR_FMT_FASTQ);
    s.fmt = fmt;
    if (nbytes && fmt == 0) {  // neither '>' nor '@': needletail refuses the stream
        s.host_error = true;
        s.host_error_bits = SBR_ST_INVALID_START;
        s.in_flight = true;
        s.nbytes_eff = 0;
        return SBR_OK;
    }
    if (final_batch && nbytes) {
        // end of stream: supply a missing final newline; FASTQ tolerates up to three trailing blank
        // lines (needletail), which are dropped here; more are left in place and surface as a bad record
        uint64_t e = nbytes;
        while (e > 0 && (h[e - 1] == '\n' || h[e - 1] == '\r')) --e;
        uint64_t newlines = 0;
        for (uint64_t q = e; q < nbytes; ++q) newlines += h[q] == '\n';
        if (newlines == 0) {
            s.supplied_nl = true;
            if (e == nbytes) { h[nbytes] = '\n'; eff = nbytes + 1; }  // slot buffers have 64 spare bytes
            else { h[e] = '\n'; eff = e + 1; }                        // trailing lone CRs
        } else if (fmt == SBR_FMT_FASTQ && newlines <= 4 && e > 0) {
            // keep the record's own line end, drop <= 3 blank lines
            if (h[e] == '\n') eff = e + 1;
            else if (h[e] == '\r' && e + 1 < nbytes && h[e + 1] == '\n') eff = e + 2;
        }
    }
    s.nbytes_eff = (uint32_t)eff;
    if (eff) {
        memset(h + eff, 0, 16);
        uint64_t copy = (eff + 15) & ~15ull;
        CU(c, cudaMemcpyAsync(s.d_text, h, copy, cudaMemcpyHostToDevice, s.stream));
        rc = launch_pipeline(c, s, s.d_text, (uint32_t)eff, fmt, final_batch, s.stream, false);
        if (rc) return rc;
        CU(c, cudaMemcpyAsync(s.h_hdr, s.d_hdr, sizeof(BatchHeader), cudaMemcpyDeviceToHost, s.stream));
        CU(c, cudaMemcpyAsync(s.h_bin_count, s.d_bin_count, c->nbins * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        CU(c, cudaMemcpyAsync(s.h_bin_start, s.d_bin_start, (c->nbins + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                              s.stream));
        CU(c, cudaEventRecord(s.hdr_ready, s.stream));
    }
    s.in_flight = true;
    return SBR_OK;
}

extern "C" int sbr_demux_device(sbr_ctx* c, uint32_t slot, const uint8_t* d_text, uint64_t nbytes, int final_batch,
                                void* stream) {
    if (!c || slot >= c->slots.size() || !d_text) return fail(c, SBR_E_INVALID_ARG, "bad argument");
    if (nbytes == 0 || nbytes > SCAN_MAX_BATCH) return fail(c, SBR_E_INVALID_ARG, "nbytes must be in (0, 1 GiB]");
    if (scan_tiles((uint32_t)nbytes) > c->max_tiles) return fail(c, SBR_E_INVALID_ARG, "nbytes exceeds batch_bytes");
    if ((uintptr_t)d_text & 15u) return fail(c, SBR_E_INVALID_ARG, "d_text must be 16-byte aligned");
    if (c->cfg.fmt == SBR_FMT_AUTO) return fail(c, SBR_E_INVALID_ARG, "device-resident batches need an explicit fmt");
    Slot& s = c->slots[slot];
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : s.stream;
    s.batch_seq = 0;
    s.nbytes_in = nbytes;
    s.nbytes_eff = (uint32_t)nbytes;
    s.final_batch = final_batch;
    s.fmt = c->cfg.fmt;
    s.host_error = false;
    s.last_stream = c->overlap ? s.aux : st;
    s.via_submit = false;
    s.supplied_nl = false;
    int rc = launch_pipeline(c, s, d_text, (uint32_t)nbytes, c->cfg.fmt, final_batch, st, c->overlap);
    if (rc) return rc;
    s.in_flight = true;
    return SBR_OK;
}

extern "C" int sbr_device_join(sbr_ctx* c, void* stream) {
    if (!c) return fail(c, SBR_E_INVALID_ARG, "bad argument");
    CU(c, cudaSetDevice(c->device));
    for (auto& s : c->slots)
        if (s.mp_pending) CU(c, cudaStreamWaitEvent((cudaStream_t)stream, s.mp_done, 0));
    return SBR_OK;
}

extern "C" int sbr_scan_stats(sbr_ctx* c, uint64_t out[2], int reset) {
    if (!c || !out) return fail(c, SBR_E_INVALID_ARG, "bad argument");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    uint32_t h[4] = {0, 0, 0, 0};
    CU(c, cudaMemcpy(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost));
    out[0] = h[0];
    out[1] = h[1];
    if (reset) CU(c, cudaMemset(c->d_stats, 0, sizeof h));
    return SBR_OK;
}

extern "C" int sbr_scan_geometry(const sbr_ctx* c, uint32_t fmt, uint32_t* tile_bytes, uint32_t* regions) {
    if (!c || (fmt != SBR_FMT_FASTA && fmt != SBR_FMT_FASTQ)) return SBR_E_INVALID_ARG;
    if (tile_bytes) *tile_bytes = scan_fast_tile_bytes();
    if (regions) *regions = (uint32_t)(fmt == SBR_FMT_FASTA ? c->fast_grid_a : c->fast_grid_q);
    return SBR_OK;
}

extern "C" int sbr_slot_device_result(sbr_ctx* c, uint32_t slot, sbr_result* r) {
    if (!c || slot >= c->slots.size() || !r) return fail(c, SBR_E_INVALID_ARG, "bad argument");
    Slot& s = c->slots[slot];
    memset(r, 0, sizeof *r);
    r->n_bins = c->nbins;
    r->rec_off = s.d_rec_off;
    r->assign = s.d_assign;
    r->bin_count = s.d_bin_count;
    r->bin_start = s.d_bin_start;
    r->order = s.d_order;
    r->rec_key = s.d_key;
    return SBR_OK;
}

extern "C" int sbr_wait(sbr_ctx* c, uint32_t slot, sbr_result* r) {
    if (!c || slot >= c->slots.size() || !r) return fail(c, SBR_E_INVALID_ARG, "bad argument");
    Slot& s = c->slots[slot];
    if (!s.in_flight) return fail(c, SBR_E_INVALID_ARG, "slot %u has nothing in flight", slot);
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_host_mirrors(c, s);
    if (rc) return rc;
    memset(r, 0, sizeof *r);
    r->batch_seq = s.batch_seq;
    r->n_bins = c->nbins;
    r->fmt = s.fmt;
    r->rec_off = s.h_rec_off;
    r->assign = s.h_assign;
    r->bin_count = s.h_bin_count;
    r->bin_start = s.h_bin_start;
    r->order = s.h_order;
    r->rec_key = nullptr;
    s.in_flight = false;
    if (s.host_error || s.nbytes_eff == 0) {
        memset(s.h_bin_count, 0, c->nbins * sizeof(uint32_t));
        memset(s.h_bin_start, 0, (c->nbins + 1) * sizeof(uint32_t));
        s.h_rec_off[0] = 0;
        r->status_flags = s.host_error_bits;
        r->first_bad_flags = s.host_error_bits;
        return SBR_OK;
    }
    cudaStream_t st = s.last_stream;
    if (s.via_submit) {
        CU(c, cudaEventSynchronize(s.hdr_ready));
    } else {  // device-resident batch: fetch the small results now
        CU(c, cudaMemcpyAsync(s.h_hdr, s.d_hdr, sizeof(BatchHeader), cudaMemcpyDeviceToHost, st));
        CU(c, cudaMemcpyAsync(s.h_bin_count, s.d_bin_count, c->nbins * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(c, cudaMemcpyAsync(s.h_bin_start, s.d_bin_start, (c->nbins + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
    }
    if (s.h_hdr->spec_fail && !c->exact_only) {  // rare: the optimistic scan gave up, the exact one takes the batch
        rc = rerun_exact(c, s, st);
        if (rc) return rc;
        CU(c, cudaMemcpyAsync(s.h_hdr, s.d_hdr, sizeof(BatchHeader), cudaMemcpyDeviceToHost, st));
        CU(c, cudaMemcpyAsync(s.h_bin_count, s.d_bin_count, c->nbins * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(c, cudaMemcpyAsync(s.h_bin_start, s.d_bin_start, (c->nbins + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
    }
    const BatchHeader& H = *s.h_hdr;
    const uint32_t n = H.n_records;
    r->n_records = n;
    r->fmt = H.fmt ? H.fmt : s.fmt;
    r->scan_path = H.mode;
    // second phase: the per-record arrays at their exact size
    CU(c, cudaMemcpyAsync(s.h_rec_off, s.d_rec_off, ((size_t)n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (n) {
        CU(c, cudaMemcpyAsync(s.h_assign, s.d_assign, (size_t)n * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
        CU(c, cudaMemcpyAsync(s.h_order, s.d_order, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    uint32_t flags = H.status & (SBR_ST_REC_OVERFLOW | SBR_ST_NONCANONICAL);
    const bool patched_end = s.supplied_nl && n && true;
    if (patched_end) flags |= SBR_ST_NONCANONICAL;
    if ((flags & SBR_ST_NONCANONICAL) && n) {
        if (!s.h_key) CU(c, cudaMallocHost(&s.h_key, ((size_t)c->max_records + 4) * sizeof(uint64_t)));
        CU(c, cudaMemcpyAsync(s.h_key, s.d_key, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        r->rec_key = s.h_key;
    }
    CU(c, cudaStreamSynchronize(st));
    const uint64_t consumed_eff = s.h_rec_off[n];  // in the (possibly newline-patched) device text
    if (patched_end && s.h_rec_off[n] > s.nbytes_in) {
        // the last record ends at the caller's last byte; its canonical form gains the newline
        s.h_rec_off[n] = (uint32_t)s.nbytes_in;
        s.h_key[n - 1] |= KEY_NONCANON;
    }
    r->consumed = s.h_rec_off[n];
    if (H.first_bad_inv) {
        unsigned long long packed = ~H.first_bad_inv;
        uint32_t rec = (uint32_t)(packed >> 8), bits = (uint32_t)(packed & 0xFF);
        if (rec < n) {  // violations attributed to the unconsumed tail are not errors of this batch
            flags |= bits;
            r->first_bad_record = rec;
            r->first_bad_flags = bits;
        }
    }
    if (s.final_batch && consumed_eff < s.nbytes_eff && !(flags & SBR_ST_REC_OVERFLOW)) {
        // the stream ends inside a record (blank tails were removed at submit time)
        if (!(flags & SBR_ST_ERROR_MASK)) { r->first_bad_record = n; r->first_bad_flags = SBR_ST_TRUNCATED; }
        flags |= SBR_ST_TRUNCATED;
    }
    if (r->consumed > s.nbytes_in) r->consumed = s.nbytes_in;  // the supplied final newline is not the caller's byte
    r->status_flags = flags;
    return SBR_OK;
}

extern "C" int sbr_debug_regions(sbr_ctx* c, uint32_t slot, uint32_t* host_out, uint32_t max_regions) {
    if (!c || slot >= c->slots.size() || !host_out) return fail(c, SBR_E_INVALID_ARG, "bad argument");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    uint32_t n = max_regions < (uint32_t)MAX_REGIONS ? max_regions : (uint32_t)MAX_REGIONS;
    CU(c, cudaMemcpy(host_out, c->slots[slot].d_info, (size_t)n * sizeof(RegionInfo), cudaMemcpyDeviceToHost));
    return (int)n;
}

extern "C" uint64_t sbr_lut_entries(const sbr_ctx* c) { return c ? c->lut_entries : 0; }

extern "C" int sbr_lut_copy(sbr_ctx* c, uint16_t* host_out, uint64_t n_entries) {
    if (!c || !host_out || !c->d_lut || n_entries > c->lut_entries) return fail(c, SBR_E_INVALID_ARG, "no LUT / bad size");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpy(host_out, c->d_lut, n_entries * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    return SBR_OK;
}

extern "C" int sbr_timing_collect(sbr_ctx* c, double ms[3], uint64_t* batches, uint64_t* launches) {
    if (!c || !ms) return fail(c, SBR_E_INVALID_ARG, "bad argument");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    ms[0] = ms[1] = ms[2] = 0.0;
    for (int i = 0; i < c->ev_used; ++i) {
        cudaEvent_t* ev = &c->ev[(size_t)i * EV_PER];
        const int from[3] = {0, 4, 2}, to[3] = {1, 2, 3};
        for (int k = 0; k < 3; ++k) {
            float t = 0.f;
            CU(c, cudaEventElapsedTime(&t, ev[from[k]], ev[to[k]]));
            ms[k] += t;
        }
    }
    if (batches) *batches = (uint64_t)c->ev_used;
    if (launches) *launches = c->launches;
    c->ev_used = 0;
    c->launches = 0;
    return SBR_OK;
}

// ---- plain memory helpers ----------------------------------------------------------------------
extern "C" int sbr_dev_alloc(int device, uint64_t nbytes, void** d_ptr) {
    if (!d_ptr) return SBR_E_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, SBR_E_NO_DEVICE, "cudaSetDevice(%d) failed", device);
    cudaError_t e = cudaMalloc(d_ptr, nbytes);
    return e == cudaSuccess ? SBR_OK : fail(nullptr, SBR_E_NOMEM, "cudaMalloc(%llu): %s", (unsigned long long)nbytes,
                                            cudaGetErrorString(e));
}
extern "C" int sbr_dev_free(int device, void* d_ptr) {
    if (cudaSetDevice(device) != cudaSuccess) return SBR_E_NO_DEVICE;
    return cudaFree(d_ptr) == cudaSuccess ? SBR_OK : SBR_E_CUDA;
}
extern "C" int sbr_dev_upload(int device, void* d_dst, const void* h_src, uint64_t nbytes) {
    if (cudaSetDevice(device) != cudaSuccess) return SBR_E_NO_DEVICE;
    return cudaMemcpy(d_dst, h_src, nbytes, cudaMemcpyHostToDevice) == cudaSuccess ? SBR_OK : SBR_E_CUDA;
}
extern "C" int sbr_dev_download(int device, void* h_dst, const void* d_src, uint64_t nbytes) {
    if (cudaSetDevice(device) != cudaSuccess) return SBR_E_NO_DEVICE;
    return cudaMemcpy(h_dst, d_src, nbytes, cudaMemcpyDeviceToHost) == cudaSuccess ? SBR_OK : SBR_E_CUDA;
}
extern "C" int sbr_host_alloc_pinned(uint64_t nbytes, void** h_ptr) {
    if (!h_ptr) return SBR_E_INVALID_ARG;
    return cudaMallocHost(h_ptr, nbytes) == cudaSuccess ? SBR_OK : SBR_E_NOMEM;
}
extern "C" int sbr_host_free_pinned(void* h_ptr) { return cudaFreeHost(h_ptr) == cudaSuccess ? SBR_OK : SBR_E_CUDA; }
