// k_match.cu -- kernel 2: barcode match with <= m Hamming mismatches, first match in table order.
//
// Replaces  barcodes.iter().find(|bc| bc_cmp(bc, &seq[..bc_len], mismatch))  (src/demux.rs:52-54,
// 103, 116) and bc_cmp itself (src/utils.rs:115-123):
//   * acceptance: |{i < min(len_b, bc_len) : barcode[i] != read[i]}| <= m, byte inequality;
//   * tie-break : the FIRST row in barcode-file order, not the best one;
//   * miss      : bin n_barcodes ("unknown").
// Packed path: reads arrive as 8-byte keys (2-bit codes + per-base invalid bits) written by the scan
// kernel; barcodes are 2-bit packed in shared memory; mismatches = popc of the pair-collapsed XOR.
// A read base that is not A/C/G/T differs from every (ACGT) barcode base, so its invalid bit is OR-ed
// into the mismatch word.  For 4^bc_len <= 2^24 the table walk is folded once per context into a
// first-match lookup table (built by the same XOR+popc routine over all 4^k keys) and the per-read
// work becomes one L1/L2 gather; reads with invalid bases still take the table walk.
// Generic path (non-ACGT barcode bytes or bc_len > 16): byte-exact compare on the raw text.
#include "scan_common.cuh"

namespace sbr {

namespace {

constexpr int MATCH_THREADS = 256;

__device__ __forceinline__ uint32_t spread16(uint32_t v) {  // bit i -> bit 2i
    v = (v | (v << 8)) & 0x00FF00FFu;
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

// first row whose collapsed XOR (+ invalid positions) has <= m bits inside the row's care mask
__device__ __forceinline__ uint32_t first_match(const uint2* __restrict__ tab, uint32_t n, uint32_t m, uint32_t codes,
                                                uint32_t inval_spread) {
    for (uint32_t b = 0; b < n; ++b) {
        uint2 row = tab[b];  // x = codes, y = care mask (even bits of compared positions)
        uint32_t x = codes ^ row.x;
        uint32_t t = ((x | (x >> 1)) | inval_spread) & row.y;
        if ((uint32_t)__popc(t) <= m) return b;
    }
    return n;
}

__global__ void __launch_bounds__(MATCH_THREADS) k_build_lut(const uint2* __restrict__ g_tab, uint32_t n, uint32_t m,
                                                             uint32_t entries, uint16_t* __restrict__ lut) {
    extern __shared__ uint2 s_tab[];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_tab[i] = g_tab[i];
    __syncthreads();
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < entries; idx += gridDim.x * blockDim.x)
        lut[idx] = (uint16_t)first_match(s_tab, n, m, idx, 0u);
}

// Where record g of the batch lives.  mode 1 (fast scan): per-region segments, found by a binary search
// over the regions' exclusive record prefix; mode 0 (look-back scan): dense arrays.
struct RecSource {
    const uint64_t* key_dense;
    const uint64_t* seg_key;
    const uint32_t* seg_rec_off;
    const uint32_t* region_base;  // [n_regions + 1]
    uint32_t cap_r;
};

// records [lo, hi) of warp-slice v out of V: equal contiguous chunks, multiple of 32 (same as k_partition.cu)
__device__ __forceinline__ void slice_range(uint32_t n, uint32_t V, uint32_t v, uint32_t& lo, uint32_t& hi) {
    uint32_t per = (n + V - 1) / V;
    per = (per + 31u) & ~31u;
    uint64_t a = (uint64_t)per * v, b = a + per;
    lo = a < n ? (uint32_t)a : n;
    hi = b < n ? (uint32_t)b : n;
}

// MODE 0: first-match LUT, 1: XOR+popc table walk, 2: byte-exact compare on the raw text.
// Every warp owns a contiguous slice of the batch's records (the slices of kernel 3): it reads the
// scan words (compacting the fast scan's per-region segments into the dense rec_off array on the way),
// writes the assignments and counts its bins in shared memory -> cta_cnt[bin][cta].
template <int MODE>
__global__ void __launch_bounds__(MATCH_THREADS) k_match(RecSource src, const BatchHeader* __restrict__ hdr,
                                                         const uint2* __restrict__ g_tab, uint32_t n, uint32_t m,
                                                         uint32_t code_mask, const uint16_t* __restrict__ lut,
                                                         const uint8_t* __restrict__ text, uint32_t nbytes,
                                                         const uint8_t* __restrict__ g_rows,
                                                         const uint32_t* __restrict__ g_lens, uint32_t stride, uint32_t k,
                                                         uint16_t* __restrict__ assign, uint32_t* __restrict__ rec_off,
                                                         uint64_t* __restrict__ key_out, uint32_t* __restrict__ cta_cnt,
                                                         uint32_t nbins) {
    extern __shared__ __align__(16) uint8_t s_mem[];
    // layout: [warps][nbins] histogram | region_base copy | table
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_mem);
    constexpr int WARPS = MATCH_THREADS / 32;
    uint32_t* s_base = s_hist + (size_t)WARPS * nbins;
    const uint32_t mode = hdr->mode, nreg = mode ? hdr->n_regions : 0u;
    uint8_t* s_tab_raw = reinterpret_cast<uint8_t*>(s_base + MAX_REGIONS + 2);  // even word count: uint2 rows stay 8-byte aligned
    uint2* s_tab = reinterpret_cast<uint2*>(s_tab_raw);
    uint8_t* s_rows = s_tab_raw;
    uint32_t* s_lens = reinterpret_cast<uint32_t*>(s_tab_raw + (((size_t)n * stride + 3) & ~(size_t)3));
    if (MODE == 2) {
        for (uint32_t i = threadIdx.x; i < n * stride; i += blockDim.x) s_rows[i] = g_rows[i];
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_lens[i] = min(g_lens[i], k);  // zip stops at seq[..bc_len]
    } else {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_tab[i] = g_tab[i];
    }
    for (uint32_t i = threadIdx.x; i <= nreg; i += blockDim.x) s_base[i] = src.region_base[i];
    for (uint32_t i = threadIdx.x; i < WARPS * nbins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const uint32_t nrec = hdr->n_records;
    const bool fasta = hdr->fmt == SBR_FMT_FASTA;
    // the host asks for the per-record scan words only when some record needs re-serialisation
    const bool want_keys = mode && (hdr->status & SBR_ST_NONCANONICAL);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t V = gridDim.x * WARPS, v = blockIdx.x * WARPS + warp;
    uint32_t* mine = s_hist + (size_t)warp * nbins;
    uint32_t lo, hi;
    slice_range(nrec, V, v, lo, hi);
    // four records per lane and step: all loads of a step are issued before the first use
    constexpr int U = 4;
    uint32_t reg_lo = 1, reg_hi = 0;  // record range [reg_lo, reg_hi) of the region found last, and where its segment starts
    size_t reg_at = 0;
    for (uint32_t g0 = lo; g0 < hi; g0 += 32 * U) {
        uint64_t kk[U];
        uint32_t ro[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t g = g0 + u * 32 + lane;
            kk[u] = KEY_SHORT;
            ro[u] = 0;
            if (g < hi) {
                if (mode) {
                    if (g < reg_lo || g >= reg_hi) {  // regions hold thousands of records: rarely taken
                        uint32_t a = 0, b = nreg;     // largest c with s_base[c] <= g
                        while (b - a > 1) {
                            uint32_t mid = (a + b) >> 1;
                            if (s_base[mid] <= g) a = mid; else b = mid;
                        }
                        reg_lo = s_base[a];
                        reg_hi = s_base[a + 1];
                        reg_at = (size_t)a * src.cap_r;
                    }
                    const size_t at = reg_at + (g - reg_lo);
                    kk[u] = ld_stream_u64(src.seg_key + at);
                    ro[u] = src.seg_rec_off[at];
                } else {
                    kk[u] = src.key_dense[g];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t g = g0 + u * 32 + lane;
            const bool live = g < hi;
            uint32_t hit = n;
            const bool usable = live && !(kk[u] & KEY_SHORT);  // short reads are reported by the scan (the reference panics)
            if (MODE == 2) {
                if (usable) {
                    uint8_t pre[64];
                    uint32_t q = (uint32_t)kk[u], got = 0;
                    while (got < k && q < nbytes) {  // FASTA sequences may wrap lines inside the window
                        uint8_t c = text[q++];
                        if (fasta && (c == '\n' || (c == '\r' && q < nbytes && text[q] == '\n'))) continue;
                        pre[got++] = c;
                    }
                    for (uint32_t b = 0; b < n && hit == n; ++b) {
                        const uint8_t* row = s_rows + (size_t)b * stride;
                        uint32_t mis = 0;
                        for (uint32_t j = 0; j < s_lens[b]; ++j) mis += (row[j] != pre[j]);
                        if (mis <= m) hit = b;
                    }
                }
            } else {
                const uint32_t codes = (uint32_t)kk[u] & code_mask;
                const uint32_t inval = (uint32_t)(kk[u] >> 32) & 0xFFFFu;
                if (MODE == 1) {
                    if (usable) hit = first_match(s_tab, n, m, codes, spread16(inval));
                } else {
                    if (usable && inval == 0) hit = __ldg(lut + codes);
                    // reads with a non-ACGT base in the window (rare) cannot use the LUT: the warp walks the
                    // table together, 32 rows per step, one such read at a time
                    uint32_t todo = __ballot_sync(0xFFFFFFFFu, usable && inval != 0);
                    while (todo) {
                        const int srcl = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const uint32_t cs = __shfl_sync(0xFFFFFFFFu, codes, srcl);
                        const uint32_t is = spread16(__shfl_sync(0xFFFFFFFFu, inval, srcl));
                        uint32_t found = n;
                        for (uint32_t b0 = 0; b0 < n; b0 += 32) {
                            const uint32_t b = b0 + lane;
                            bool ok = false;
                            if (b < n) {
                                const uint2 row = s_tab[b];
                                const uint32_t x = cs ^ row.x;
                                ok = (uint32_t)__popc(((x | (x >> 1)) | is) & row.y) <= m;
                            }
                            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
                            if (bal) { found = b0 + __ffs(bal) - 1; break; }
                        }
                        if (lane == srcl) hit = found;
                    }
                }
            }
            if (live) {
                assign[g] = (uint16_t)hit;
                if (mode) {
                    rec_off[g] = ro[u];
                    if (want_keys) key_out[g] = kk[u];
                }
                atomicAdd(&mine[hit], 1u);
            }
        }
    }
    // kernel 3's input: per bin, how many records the whole CTA holds
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) run += s_hist[(size_t)w * nbins + b];
        cta_cnt[(size_t)b * gridDim.x + blockIdx.x] = run;
    }
}

}  // namespace

cudaError_t match_build_lut(const uint2* d_tab, uint32_t n, uint32_t m, uint32_t k, uint16_t* d_lut,
                            cudaStream_t stream) {
    uint32_t entries = 1u << (2 * k);
    int grid = (int)min((entries + MATCH_THREADS - 1) / MATCH_THREADS, 148u * 8u);
    k_build_lut<<<grid, MATCH_THREADS, n * sizeof(uint2), stream>>>(d_tab, n, m, entries, d_lut);
    return cudaGetLastError();
}

struct MatchLaunch {
    // sources
    const uint64_t* key_dense; const uint64_t* seg_key; const uint32_t* seg_rec_off; const uint32_t* region_base;
    uint32_t cap_r;
    const BatchHeader* hdr;
    // table
    const uint2* tab; const uint16_t* lut; const uint8_t* rows; const uint32_t* lens;
    uint32_t n, m, k, stride;
    const uint8_t* text; uint32_t nbytes;
    // outputs
    uint16_t* assign; uint32_t* rec_off; uint64_t* key_out; uint32_t* cta_cnt; uint32_t nbins;
    int grid;  // = the partition plan's grid: the histogram slices are kernel 3's slices
};

size_t match_smem(uint32_t n, uint32_t stride, uint32_t nbins, bool generic) {
    size_t tab = generic ? ((((size_t)n * stride + 3) & ~(size_t)3) + n * 4u) : (size_t)n * sizeof(uint2);
    return (size_t)(MATCH_THREADS / 32) * nbins * 4u + (MAX_REGIONS + 2) * 4u + tab + 16;
}

cudaError_t match_prepare() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_match<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_match<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_match<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_build_lut, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
}

cudaError_t match_launch(const MatchLaunch& L, cudaStream_t stream) {
    RecSource src{L.key_dense, L.seg_key, L.seg_rec_off, L.region_base, L.cap_r};
    const bool generic = L.tab == nullptr;
    uint32_t code_mask = L.k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * L.k)) - 1u);
    size_t smem = match_smem(L.n, L.stride, L.nbins, generic);
    if (generic)
        k_match<2><<<L.grid, MATCH_THREADS, smem, stream>>>(src, L.hdr, nullptr, L.n, L.m, 0u, nullptr, L.text, L.nbytes, L.rows,
                                                            L.lens, L.stride, L.k, L.assign, L.rec_off, L.key_out, L.cta_cnt, L.nbins);
    else if (L.lut)
        k_match<0><<<L.grid, MATCH_THREADS, smem, stream>>>(src, L.hdr, L.tab, L.n, L.m, code_mask, L.lut, nullptr, 0u, nullptr,
                                                            nullptr, 0u, L.k, L.assign, L.rec_off, L.key_out, L.cta_cnt, L.nbins);
    else
        k_match<1><<<L.grid, MATCH_THREADS, smem, stream>>>(src, L.hdr, L.tab, L.n, L.m, code_mask, nullptr, nullptr, 0u, nullptr,
                                                            nullptr, 0u, L.k, L.assign, L.rec_off, L.key_out, L.cta_cnt, L.nbins);
    return cudaGetLastError();
}

}  // namespace sbr
