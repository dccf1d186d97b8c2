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
#include "sbr_common.cuh"

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

template <bool USE_LUT>
__global__ void __launch_bounds__(MATCH_THREADS) k_match_packed(const uint64_t* __restrict__ key,
                                                                const BatchHeader* __restrict__ hdr,
                                                                const uint2* __restrict__ g_tab, uint32_t n, uint32_t m,
                                                                uint32_t code_mask, const uint16_t* __restrict__ lut,
                                                                uint16_t* __restrict__ assign) {
    extern __shared__ uint2 s_tab[];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_tab[i] = g_tab[i];
    __syncthreads();
    const uint32_t nrec = hdr->n_records;
    auto one = [&](uint64_t k) -> uint32_t {
        uint32_t codes = (uint32_t)k & code_mask;
        uint32_t inval = (uint32_t)(k >> 32) & 0xFFFFu;
        if (k & KEY_SHORT) return n;  // reported by the scan; the host aborts like the reference's panic
        if (USE_LUT && inval == 0) return __ldg(lut + codes);
        return first_match(s_tab, n, m, codes, spread16(inval));
    };
    // four reads per thread: 2 x 16-byte key loads, one 8-byte store
    const uint32_t quads = nrec >> 2;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
        const uint4* kp = reinterpret_cast<const uint4*>(key + 4ull * q);
        uint4 a = __ldg(kp), b = __ldg(kp + 1);
        uint32_t r0 = one(((uint64_t)a.y << 32) | a.x), r1 = one(((uint64_t)a.w << 32) | a.z);
        uint32_t r2 = one(((uint64_t)b.y << 32) | b.x), r3 = one(((uint64_t)b.w << 32) | b.z);
        reinterpret_cast<uint2*>(assign)[q] = make_uint2(r0 | (r1 << 16), r2 | (r3 << 16));
    }
    if (blockIdx.x == 0) {
        uint32_t i = (quads << 2) + threadIdx.x;
        if (i < nrec) assign[i] = (uint16_t)one(key[i]);
    }
}

// byte-exact path: key[31:0] = batch offset of the first sequence byte
__global__ void __launch_bounds__(MATCH_THREADS) k_match_generic(const uint8_t* __restrict__ text, uint32_t nbytes,
                                                                 const uint64_t* __restrict__ key,
                                                                 const BatchHeader* __restrict__ hdr,
                                                                 const uint8_t* __restrict__ g_rows,
                                                                 const uint32_t* __restrict__ g_lens, uint32_t n,
                                                                 uint32_t stride, uint32_t k, uint32_t m,
                                                                 uint16_t* __restrict__ assign) {
    extern __shared__ uint8_t s_raw[];
    uint8_t* rows = s_raw;
    uint32_t* lens = reinterpret_cast<uint32_t*>(s_raw + (((size_t)n * stride + 3) & ~(size_t)3));
    for (uint32_t i = threadIdx.x; i < n * stride; i += blockDim.x) rows[i] = g_rows[i];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) lens[i] = min(g_lens[i], k);  // zip stops at seq[..bc_len]
    __syncthreads();
    const uint32_t nrec = hdr->n_records;
    const bool fasta = hdr->fmt == SBR_FMT_FASTA;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += gridDim.x * blockDim.x) {
        uint64_t kk = key[i];
        uint32_t hit = n;
        if (!(kk & KEY_SHORT)) {
            uint8_t pre[64];
            uint32_t q = (uint32_t)kk, got = 0;
            while (got < k && q < nbytes) {  // FASTA sequences may wrap lines inside the window
                uint8_t c = text[q++];
                if (fasta && (c == '\n' || (c == '\r' && q < nbytes && text[q] == '\n'))) continue;
                pre[got++] = c;
            }
            for (uint32_t b = 0; b < n && hit == n; ++b) {
                const uint8_t* row = rows + (size_t)b * stride;
                uint32_t mis = 0;
                for (uint32_t j = 0; j < lens[b]; ++j) mis += (row[j] != pre[j]);
                if (mis <= m) hit = b;
            }
        }
        assign[i] = (uint16_t)hit;
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

cudaError_t match_launch_packed(const uint64_t* d_key, const BatchHeader* d_hdr, const uint2* d_tab, uint32_t n,
                                uint32_t m, uint32_t k, const uint16_t* d_lut, uint16_t* d_assign, uint32_t max_records,
                                int sms, cudaStream_t stream) {
    uint32_t code_mask = k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    uint32_t want = (max_records / 4 + MATCH_THREADS - 1) / MATCH_THREADS;
    int grid = (int)max(1u, min(want, (uint32_t)sms * 8u));
    size_t smem = n * sizeof(uint2);
    if (d_lut)
        k_match_packed<true><<<grid, MATCH_THREADS, smem, stream>>>(d_key, d_hdr, d_tab, n, m, code_mask, d_lut, d_assign);
    else
        k_match_packed<false><<<grid, MATCH_THREADS, smem, stream>>>(d_key, d_hdr, d_tab, n, m, code_mask, nullptr,
                                                                      d_assign);
    return cudaGetLastError();
}

size_t match_generic_smem(uint32_t n, uint32_t stride) { return (((size_t)n * stride + 3) & ~(size_t)3) + n * 4u; }

cudaError_t match_prepare() {
    cudaError_t e = cudaFuncSetAttribute(k_match_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_match_packed<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_match_packed<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_build_lut, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
}

cudaError_t match_launch_generic(const uint8_t* d_text, uint32_t nbytes, const uint64_t* d_key, const BatchHeader* d_hdr,
                                 const uint8_t* d_rows, const uint32_t* d_lens, uint32_t n, uint32_t stride, uint32_t k,
                                 uint32_t m, uint16_t* d_assign, uint32_t max_records, int sms, cudaStream_t stream) {
    uint32_t want = (max_records + MATCH_THREADS - 1) / MATCH_THREADS;
    int grid = (int)max(1u, min(want, (uint32_t)sms * 4u));
    k_match_generic<<<grid, MATCH_THREADS, match_generic_smem(n, stride), stream>>>(d_text, nbytes, d_key, d_hdr, d_rows,
                                                                                    d_lens, n, stride, k, m, d_assign);
    return cudaGetLastError();
}

}  // namespace sbr
