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
#include <cooperative_groups.h>

#include "scan_common.cuh"

namespace cg = cooperative_groups;

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

struct MatchArgs {
    RecSource src;
    const BatchHeader* hdr;
    const uint2* g_tab; uint32_t n, m, code_mask;
    const uint16_t* lut;
    const uint16_t* lut1;  // the same table for mismatch - 1 (NULL when m == 0 or the rows differ in length)
    uint32_t uniform;      // every row is compared over the whole window
    const uint8_t* text; uint32_t nbytes;
    const uint8_t* g_rows; const uint32_t* g_lens; uint32_t stride, k;
    uint16_t* assign; uint32_t* rec_off; uint64_t* key_out; uint32_t* cta_cnt; uint32_t nbins;
    // FUSED only: kernel 3's outputs
    uint32_t nbits; uint32_t* bin_count; uint32_t* bin_start; uint32_t* order;
    BatchHeader* zero_next;  // the slot's other header: cleared here for the next batch (saves a memset per batch)
    uint32_t force_keys;
};

// lanes holding the same bin as this lane, from one ballot per bit of the bin index (same as k_partition.cu)
__device__ __forceinline__ uint32_t same_bin_lanes(uint32_t bin, uint32_t nbits, bool live) {
    uint32_t peers = __ballot_sync(0xFFFFFFFFu, live);
    if (!live) peers = ~peers;
    for (uint32_t k = 0; k < nbits; ++k) {
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, (bin >> k) & 1u);
        peers &= ((bin >> k) & 1u) ? b : ~b;
    }
    return peers;
}

// MODE 0: first-match LUT, 1: XOR+popc table walk, 2: byte-exact compare on the raw text.
// Every warp owns a contiguous slice of the batch's records (the slices of kernel 3): it reads the
// scan words (compacting the fast scan's per-region segments into the dense rec_off array on the way),
// writes the assignments and counts its bins in shared memory -> cta_cnt[bin][cta].
// FUSED (cooperative launch, every CTA resident): kernel 3 follows in the same launch -- grid barrier, the per-bin
// scans over CTAs spread over the grid, grid barrier, cursors from the counts still sitting in shared memory, scatter.
// LEAN (packed modes only): the shared-memory diet for big tables that have to fit beside the record scan (overlapped
// mode, one CTA per SM next to four of the scan's): per-warp bin counters are 16 bits wide, cursors are kept as
// {32-bit CTA base per bin} + {16-bit offset per warp and bin} (the context only selects this kernel when a whole CTA's
// slices hold fewer than 65536 records), and the region prefix and the barcode table are read from global memory instead
// of a shared copy.
template <int MODE, bool FUSED, bool LEAN>
__global__ void __launch_bounds__(MATCH_THREADS) k_match(const MatchArgs A) {
    const RecSource& src = A.src;
    const BatchHeader* __restrict__ hdr = A.hdr;
    const uint2* __restrict__ g_tab = A.g_tab;
    const uint32_t n = A.n, m = A.m, code_mask = A.code_mask, nbytes = A.nbytes, stride = A.stride, k = A.k, nbins = A.nbins;
    const uint16_t* __restrict__ lut = A.lut;
    const uint8_t* __restrict__ text = A.text;
    const uint8_t* __restrict__ g_rows = A.g_rows;
    const uint32_t* __restrict__ g_lens = A.g_lens;
    uint16_t* assign = A.assign;
    uint32_t* __restrict__ rec_off = A.rec_off;
    uint64_t* __restrict__ key_out = A.key_out;
    uint32_t* cta_cnt = A.cta_cnt;
    extern __shared__ __align__(16) uint8_t s_mem[];
    // layout: [warps][nbins] histogram | region_base copy | table          (LEAN: [warps][nb2] u16 | [nbins] u32 CTA base)
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_mem);
    constexpr int WARPS = MATCH_THREADS / 32;
    const uint32_t nb2 = (nbins + 1u) & ~1u;  // LEAN: 16-bit counters per warp row (even, so that rows stay word aligned)
    uint16_t* const s_h16 = reinterpret_cast<uint16_t*>(s_mem);
    uint32_t* const s_cta_base = s_hist + (size_t)WARPS * (nb2 / 2);  // LEAN only
    uint32_t* s_base = s_hist + (size_t)WARPS * nbins;
    const uint32_t mode = hdr->mode, nreg = mode ? hdr->n_regions : 0u;
    uint8_t* s_tab_raw = reinterpret_cast<uint8_t*>(s_base + MAX_REGIONS + 2);  // even word count: uint2 rows stay 8-byte aligned
    uint2* s_tab = reinterpret_cast<uint2*>(s_tab_raw);
    uint8_t* s_rows = s_tab_raw;
    uint32_t* s_lens = reinterpret_cast<uint32_t*>(s_tab_raw + (((size_t)n * stride + 3) & ~(size_t)3));
    if (LEAN) {
        for (uint32_t i = threadIdx.x; i < WARPS * (nb2 / 2); i += blockDim.x) s_hist[i] = 0;
    } else {
        if (MODE == 2) {
            for (uint32_t i = threadIdx.x; i < n * stride; i += blockDim.x) s_rows[i] = g_rows[i];
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_lens[i] = min(g_lens[i], k);  // zip stops at seq[..bc_len]
        } else {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_tab[i] = g_tab[i];
        }
        for (uint32_t i = threadIdx.x; i <= nreg; i += blockDim.x) s_base[i] = src.region_base[i];
        for (uint32_t i = threadIdx.x; i < WARPS * nbins; i += blockDim.x) s_hist[i] = 0;
    }
    const uint32_t* const rbase = LEAN ? src.region_base : s_base;  // exclusive record prefix over the regions
    const uint2* const tab = LEAN ? g_tab : s_tab;
    __syncthreads();
    const uint32_t nrec = hdr->n_records;
    const bool fasta = hdr->fmt == SBR_FMT_FASTA;
    // the host asks for the per-record scan words only when some record needs re-serialisation
    const bool want_keys = mode && ((hdr->status & SBR_ST_NONCANONICAL) || A.force_keys);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t V = gridDim.x * WARPS, v = blockIdx.x * WARPS + warp;
    uint32_t* mine = s_hist + (size_t)warp * (LEAN ? nb2 / 2 : nbins);  // LEAN: two 16-bit counters per word
    uint16_t* const mine16 = s_h16 + (size_t)warp * nb2;
    uint32_t lo, hi;
    slice_range(nrec, V, v, lo, hi);
    // U records per lane and step: all loads of a step are issued before the first use
#ifndef SBR_MATCH_U
#define SBR_MATCH_U 4
#endif
    constexpr int U = SBR_MATCH_U;
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
                        uint32_t a = 0, b = nreg;     // largest c with rbase[c] <= g
                        while (b - a > 1) {
                            uint32_t mid = (a + b) >> 1;
                            if (rbase[mid] <= g) a = mid; else b = mid;
                        }
                        reg_lo = rbase[a];
                        reg_hi = rbase[a + 1];
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
                    if (usable) hit = first_match(tab, n, m, codes, spread16(inval));
                } else {
                    if (usable && inval == 0) hit = __ldg(lut + codes);
                    // Reads with a non-ACGT base in the window: such a base differs from every barcode base
                    // (src/utils.rs:120 compares bytes).  When all rows are compared over the whole window:
                    //   * more invalid bases than the mismatch budget -> nothing can match;
                    //   * exactly one, at position p -> a row matches iff it is within m-1 of the read on the other
                    //     positions, i.e. iff it is within m-1 of the read with SOME base X put at p (X = the row's
                    //     own base there).  The first such row is the smallest answer of the (m-1)-table over the
                    //     four substitutions.
                    // Everything else (rare) walks the table: the warp together, 32 rows per step, one read at a time.
                    bool pending = usable && inval != 0;
                    if (pending && A.uniform) {
                        const uint32_t ninv = __popc(inval);
                        if (ninv > m) pending = false;  // stays unknown
                        else if (ninv == 1 && A.lut1) {
                            const uint32_t sh = 2u * (uint32_t)(__ffs(inval) - 1);
                            const uint32_t base = codes & ~(3u << sh);
                            uint32_t r = n;
#pragma unroll
                            for (uint32_t X = 0; X < 4; ++X) r = min(r, (uint32_t)__ldg(A.lut1 + (base | (X << sh))));
                            hit = r;
                            pending = false;
                        }
                    }
                    uint32_t todo = __ballot_sync(0xFFFFFFFFu, pending);
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
                                const uint2 row = tab[b];
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
                if (LEAN) atomicAdd(&mine[hit >> 1], 1u << ((hit & 1u) * 16u));
                else atomicAdd(&mine[hit], 1u);
            }
        }
    }
    // kernel 3's input: per bin, how many records the whole CTA holds
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) run += LEAN ? (uint32_t)s_h16[(size_t)w * nb2 + b] : s_hist[(size_t)w * nbins + b];
        cta_cnt[(size_t)b * gridDim.x + blockIdx.x] = run;
    }
    if (!FUSED) return;

    // ---- kernel 3 inside the same launch -----------------------------------------------------------------------
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t s_warp[WARPS];
    const uint32_t G = gridDim.x;
    __threadfence();
    grid.sync();
    // per bin: exclusive scan of the CTA totals in place (bins dealt round-robin to the CTAs), bin totals -> bin_count
    for (uint32_t b = blockIdx.x; b < nbins; b += G) {
        uint32_t* row = cta_cnt + (size_t)b * G;
        constexpr int CH = 4;
        uint32_t carry = 0;
        for (uint32_t base = 0; base < G; base += MATCH_THREADS * CH) {
            const uint32_t i0 = base + threadIdx.x * CH;
            uint32_t x[CH], sum = 0;
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                x[j] = i0 + j < G ? __ldcg(row + i0 + j) : 0u;
                sum += x[j];
            }
            const uint32_t inc = warp_incl_scan(sum, lane);
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            uint32_t pre = carry + inc - sum, tot = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t t = s_warp[w];
                if (w < warp) pre += t;
                tot += t;
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                if (i0 + j < G) row[i0 + j] = pre;
                pre += x[j];
            }
            carry += tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) A.bin_count[b] = carry;
    }
    __threadfence();
    grid.sync();
    // bin_start = exclusive scan of bin_count (every CTA for itself; CTA 0 publishes it), then the warps' cursors:
    // bin_start + what earlier CTAs hold + what the CTA's earlier warps hold (their counts are still in s_hist)
    {
        uint32_t carry = 0;
        for (uint32_t base = 0; base < nbins; base += MATCH_THREADS) {
            const uint32_t b = base + threadIdx.x;
            const uint32_t x = b < nbins ? __ldcg(A.bin_count + b) : 0u;
            const uint32_t inc = warp_incl_scan(x, lane);
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            uint32_t pre = carry + inc - x, tot = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const uint32_t t = s_warp[w];
                if (w < warp) pre += t;
                tot += t;
            }
            if (b < nbins) {
                if (blockIdx.x == 0) A.bin_start[b] = pre;
                uint32_t run = pre + __ldcg(cta_cnt + (size_t)b * G + blockIdx.x);
                if (LEAN) {  // the CTA's base for the bin + every warp's offset inside the CTA's run (fits 16 bits)
                    s_cta_base[b] = run;
                    uint32_t rel = 0;
#pragma unroll
                    for (int w = 0; w < WARPS; ++w) {
                        const uint32_t t = s_h16[(size_t)w * nb2 + b];
                        s_h16[(size_t)w * nb2 + b] = (uint16_t)rel;
                        rel += t;
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < WARPS; ++w) {
                        const uint32_t t = s_hist[(size_t)w * nbins + b];
                        s_hist[(size_t)w * nbins + b] = run;
                        run += t;
                    }
                }
            }
            carry += tot;
            __syncthreads();
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) A.bin_start[nbins] = carry;
    }
    __syncthreads();
    // stable scatter of the warp's own slice (its assignments come back from L2)
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t g0 = lo; g0 < hi; g0 += 32 * U) {
        uint32_t bins[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = g0 + u * 32 + lane;
            bins[u] = i < hi ? (uint32_t)assign[i] : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = g0 + u * 32 + lane;
            if (g0 + u * 32 >= hi) break;  // warp-uniform
            const bool live = i < hi;
            const uint32_t bin = bins[u];
            const uint32_t peers = same_bin_lanes(bin, A.nbits, live);
            const uint32_t rank = __popc(peers & lt);
            uint32_t at = 0u;  // every lane of a group reads the group's cursor ...
            if (live) at = LEAN ? s_cta_base[bin] + (uint32_t)mine16[bin] : mine[bin];
            __syncwarp();
            if (live && rank == 0) {  // ... then its first lane advances it
                if (LEAN) mine16[bin] = (uint16_t)(mine16[bin] + __popc(peers));
                else mine[bin] = at + __popc(peers);
            }
            if (live) A.order[at + rank] = i;
            __syncwarp();
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < sizeof(BatchHeader) / 4 && A.zero_next)
        reinterpret_cast<uint32_t*>(A.zero_next)[threadIdx.x] = 0u;
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
    const uint2* tab; const uint16_t* lut; const uint16_t* lut1; uint32_t uniform; const uint8_t* rows; const uint32_t* lens;
    uint32_t n, m, k, stride;
    const uint8_t* text; uint32_t nbytes;
    // outputs
    uint16_t* assign; uint32_t* rec_off; uint64_t* key_out; uint32_t* cta_cnt; uint32_t nbins;
    int grid;  // = the partition plan's grid: the histogram slices are kernel 3's slices
    // fused with kernel 3 (cooperative launch) when `fused`
    int fused; uint32_t nbits; uint32_t* bin_count; uint32_t* bin_start; uint32_t* order; BatchHeader* zero_next;
    uint32_t force_keys;  // write key_out whatever the header says (the host patched the stream's last record)
    uint32_t lean;        // the shared-memory diet (k_match<.., LEAN>): see match_smem()
};

size_t match_smem(uint32_t n, uint32_t stride, uint32_t nbins, bool generic, bool lean) {
    if (lean) return (size_t)(MATCH_THREADS / 32) * ((nbins + 1u) & ~1u) * 2u + (size_t)nbins * 4u + 16;
    size_t tab = generic ? ((((size_t)n * stride + 3) & ~(size_t)3) + n * 4u) : (size_t)n * sizeof(uint2);
    return (size_t)(MATCH_THREADS / 32) * nbins * 4u + (MAX_REGIONS + 2) * 4u + tab + 16;
}

using MatchKernel = void (*)(const MatchArgs);
static MatchKernel match_kernel(int mode, bool fused, bool lean = false) {
    if (lean && mode < 2) {  // (the byte-exact kernel keeps its rows in shared memory: no lean variant)
        switch (mode * 2 + (fused ? 1 : 0)) {
            case 0: return k_match<0, false, true>;
            case 1: return k_match<0, true, true>;
            case 2: return k_match<1, false, true>;
            default: return k_match<1, true, true>;
        }
    }
    switch (mode * 2 + (fused ? 1 : 0)) {
        case 0: return k_match<0, false, false>;
        case 1: return k_match<0, true, false>;
        case 2: return k_match<1, false, false>;
        case 3: return k_match<1, true, false>;
        case 4: return k_match<2, false, false>;
        default: return k_match<2, true, false>;
    }
}

// max_shared: the kernels will run beside the record scan (overlapped mode), whose CTAs need the SM's largest
// shared-memory split: ask for the same one, or the two could not be resident together
cudaError_t match_prepare(bool max_shared) {
    cudaError_t e;
    for (int lean = 0; lean < 2; ++lean)
        for (int mode = 0; mode < (lean ? 2 : 3); ++mode)
            for (int f = 0; f < 2; ++f) {
                const MatchKernel kf = match_kernel(mode, f != 0, lean != 0);
                if ((e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)) != cudaSuccess) return e;
                if ((e = cudaFuncSetAttribute(kf, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              max_shared ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault)) != cudaSuccess)
                    return e;
            }
    return cudaFuncSetAttribute(k_build_lut, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
}

// the largest grid of the fused kernel that is resident all at once (a cooperative launch needs that); 0 = unsupported
int match_fused_max_grid(int device, int mode, size_t smem, bool lean) {
    int coop = 0, sms = 0, per_sm = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) != cudaSuccess || !coop) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, match_kernel(mode, true, lean), MATCH_THREADS, smem) != cudaSuccess) return 0;
    return per_sm * sms;
}

cudaError_t match_launch(const MatchLaunch& L, cudaStream_t stream) {
    const bool generic = L.tab == nullptr;
    const int mode = generic ? 2 : (L.lut ? 0 : 1);
    MatchArgs A{};
    A.src = RecSource{L.key_dense, L.seg_key, L.seg_rec_off, L.region_base, L.cap_r};
    A.hdr = L.hdr;
    A.g_tab = L.tab; A.n = L.n; A.m = L.m;
    A.code_mask = generic ? 0u : (L.k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * L.k)) - 1u));
    A.lut = L.lut; A.lut1 = L.lut1; A.uniform = L.uniform;
    A.text = generic ? L.text : nullptr; A.nbytes = generic ? L.nbytes : 0u;
    A.g_rows = L.rows; A.g_lens = L.lens; A.stride = generic ? L.stride : 0u; A.k = L.k;
    A.assign = L.assign; A.rec_off = L.rec_off; A.key_out = L.key_out; A.cta_cnt = L.cta_cnt; A.nbins = L.nbins;
    A.nbits = L.nbits; A.bin_count = L.bin_count; A.bin_start = L.bin_start; A.order = L.order;
    A.zero_next = L.fused ? L.zero_next : nullptr;
    A.force_keys = L.force_keys;
    const bool lean = L.lean && !generic;
    const size_t smem = match_smem(L.n, L.stride, L.nbins, generic, lean);
    if (L.fused) {
        void* args[] = {&A};
        return cudaLaunchCooperativeKernel((const void*)match_kernel(mode, true, lean), dim3(L.grid), dim3(MATCH_THREADS), args, smem, stream);
    }
    match_kernel(mode, false, lean)<<<L.grid, MATCH_THREADS, smem, stream>>>(A);
    return cudaGetLastError();
}

}  // namespace sbr
