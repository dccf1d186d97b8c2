// k_partition.cu -- kernel 3: stable per-barcode histogram + exclusive scan + scatter.
//
// Replaces the implicit partition of the reference: every record is appended, in input order, to
// the file of the barcode it matched or to the unknown file (src/demux.rs:58-59, 62-63, 104-108,
// 117-121) and counted in nb_records (src/demux.rs:58, 104, 117).  Here the same partition is
// returned as index lists: order[bin_start[b] .. bin_start[b+1]) = the records of bin b in input
// order, bin_count[b] = how many.  Bin n_barcodes is "unknown".
//
// A stable counting sort over <= 65535 bins:
//   (hist)    fused into the match kernel: every warp owns a contiguous slice of the records and counts
//             its bins in shared memory while it writes the assignments; the CTA leaves
//             cta_cnt[bin][cta] = records of that bin in the whole CTA (8 consecutive warp slices)
//   k_binscan one block per bin: exclusive scan of cta_cnt[bin][*] in place, bin totals -> bin_count; the
//             last block to finish scans bin_count into bin_start
//   k_scatter same slices.  The CTA first recounts its warps' bins in shared memory (assign is L2-resident) and
//             turns them into cursors: bin_start + cta_cnt + what the CTA's earlier warps hold.  Equal bins
//             inside each 32-record step are ranked with one ballot per bin-index bit, so ties keep input order
//                                                                   (reads assign twice, writes 4 B/record)
#include "sbr_common.cuh"

namespace sbr {

namespace {

constexpr int PART_THREADS = 256;
constexpr int PART_WARPS = PART_THREADS / 32;

// records [slice_lo, slice_hi) of warp-slice v out of V: equal contiguous chunks, multiple of 32
__device__ __forceinline__ void slice_range(uint32_t n, uint32_t V, uint32_t v, uint32_t& lo, uint32_t& hi) {
    uint32_t per = (n + V - 1) / V;
    per = (per + 31u) & ~31u;
    uint64_t a = (uint64_t)per * v, b = a + per;
    lo = a < n ? (uint32_t)a : n;
    hi = b < n ? (uint32_t)b : n;
}

// One block per bin: in-place exclusive scan over the G CTA totals of that bin (every thread owns a contiguous
// chunk).  The last block to finish turns the bin totals into bin_start (exclusive scan over bins).
__global__ void __launch_bounds__(PART_THREADS) k_binscan(uint32_t* __restrict__ cta_cnt, uint32_t G, uint32_t nbins,
                                                          uint32_t* __restrict__ bin_count,
                                                          uint32_t* __restrict__ bin_start /* nbins + 1 */,
                                                          uint32_t* __restrict__ done_counter) {
    __shared__ uint32_t s_warp[PART_WARPS];
    __shared__ uint32_t s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* row = cta_cnt + (size_t)blockIdx.x * G;
    constexpr int CH = 8;  // entries per thread and round
    uint32_t carry = 0;
    for (uint32_t base = 0; base < G; base += PART_THREADS * CH) {
        const uint32_t i0 = base + threadIdx.x * CH;
        uint32_t x[CH], sum = 0;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            x[j] = i0 + j < G ? row[i0 + j] : 0u;
            sum += x[j];
        }
        const uint32_t inc = warp_incl_scan(sum, lane);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t pre = carry + inc - sum, tot = 0;
#pragma unroll
        for (int w = 0; w < PART_WARPS; ++w) {
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
    if (threadIdx.x == 0) {
        bin_count[blockIdx.x] = carry;
        __threadfence();
        s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    // last block: exclusive scan of bin_count -> bin_start
    __threadfence();
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) { s_carry = 0; *done_counter = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < nbins; base += PART_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < nbins ? __ldcg(bin_count + i) : 0u;
        uint32_t in2 = warp_incl_scan(x, lane);
        if (lane == 31) s_warp[warp] = in2;
        __syncthreads();
        uint32_t p2 = s_carry, t2 = 0;
#pragma unroll
        for (int w = 0; w < PART_WARPS; ++w) {
            uint32_t t = s_warp[w];
            if (w < warp) p2 += t;
            t2 += t;
        }
        if (i < nbins) bin_start[i] = p2 + in2 - x;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += t2;
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_start[nbins] = s_carry;
}

// lanes holding the same bin as this lane, from one ballot per bit of the bin index (NBITS = bits needed for nbins;
// dead lanes carry an index with a bit above NBITS set and fall into their own group through the extra ballot)
__device__ __forceinline__ uint32_t same_bin_lanes(uint32_t bin, uint32_t nbits, bool live) {
    uint32_t peers = __ballot_sync(0xFFFFFFFFu, live);
    if (!live) peers = ~peers;
    for (uint32_t k = 0; k < nbits; ++k) {
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, (bin >> k) & 1u);
        peers &= ((bin >> k) & 1u) ? b : ~b;
    }
    return peers;
}

__global__ void __launch_bounds__(PART_THREADS) k_scatter(const uint16_t* __restrict__ assign,
                                                          const BatchHeader* __restrict__ hdr, uint32_t nbins, uint32_t nbits,
                                                          const uint32_t* __restrict__ cta_cnt /* scanned over CTAs */,
                                                          const uint32_t* __restrict__ bin_start,
                                                          uint32_t* __restrict__ order, BatchHeader* zero_next) {
    extern __shared__ uint32_t s_cur[];  // [PART_WARPS][nbins]: counts first, then the running cursor per bin
    const uint32_t n = hdr->n_records;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t G = gridDim.x, V = G * PART_WARPS, v = blockIdx.x * PART_WARPS + warp;
    uint32_t* cur = s_cur + (size_t)warp * nbins;
    uint32_t lo, hi;
    slice_range(n, V, v, lo, hi);
    for (uint32_t i = threadIdx.x; i < PART_WARPS * nbins; i += PART_THREADS) s_cur[i] = 0;
    __syncthreads();
    for (uint32_t i = lo + lane; i < hi; i += 32) atomicAdd(&cur[assign[i]], 1u);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += PART_THREADS) {
        uint32_t run = bin_start[b] + cta_cnt[(size_t)b * G + blockIdx.x];
#pragma unroll
        for (int w = 0; w < PART_WARPS; ++w) {
            const uint32_t t = s_cur[(size_t)w * nbins + b];
            s_cur[(size_t)w * nbins + b] = run;
            run += t;
        }
    }
    __syncthreads();
    // the slot's other header is cleared here for the next batch (saves a memset per batch)
    if (blockIdx.x == 0 && threadIdx.x < sizeof(BatchHeader) / 4 && zero_next) reinterpret_cast<uint32_t*>(zero_next)[threadIdx.x] = 0u;
    if (lo >= hi) return;  // warp-uniform
    const uint32_t lt = (1u << lane) - 1u;
    constexpr int U = 4;  // the loads of four 32-record groups are in flight together; ranking stays in order
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
            const uint32_t peers = same_bin_lanes(bin, nbits, live);
            const uint32_t rank = __popc(peers & lt);
            // every lane of a group reads the group's cursor (a broadcast), then the group's first lane advances it
            const uint32_t base = live ? cur[bin] : 0u;
            __syncwarp();
            if (live && rank == 0) cur[bin] = base + __popc(peers);
            if (live) order[base + rank] = i;
            __syncwarp();
        }
    }
}

}  // namespace

struct PartitionPlan {
    int grid;      // blocks of the match kernel / k_scatter
    uint32_t V;    // warp slices
    size_t smem;   // dynamic shared bytes
    uint32_t nbits;  // bits of a bin index
};

PartitionPlan partition_plan(uint32_t nbins, uint32_t max_records, int sms) {
    PartitionPlan p;
    size_t smem = (size_t)PART_WARPS * nbins * sizeof(uint32_t);
    size_t fit = (200 * 1024) / (smem ? smem : 1);
    int per_sm = fit < 1 ? 1 : (fit > 8 ? 8 : (int)fit);
    uint32_t want = (max_records + PART_THREADS * 4 - 1) / (PART_THREADS * 4);  // >= 4 steps of 32 per warp
    uint32_t g = min(want, (uint32_t)(sms * per_sm));
    p.grid = (int)max(1u, g);
    p.V = (uint32_t)p.grid * PART_WARPS;
    p.smem = smem;
    p.nbits = 1;
    while ((1u << p.nbits) < nbins) ++p.nbits;
    return p;
}

cudaError_t partition_prepare(bool max_shared) {
    cudaError_t e;
    const int carve = max_shared ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault;  // see match_prepare()
    if ((e = cudaFuncSetAttribute(k_scatter, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_binscan, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}

cudaError_t partition_launch(const uint16_t* d_assign, const BatchHeader* d_hdr, uint32_t nbins,
                             uint32_t* d_cta_cnt, uint32_t* d_bin_count, uint32_t* d_bin_start, uint32_t* d_order,
                             uint32_t* d_done_counter, BatchHeader* d_zero_next, const PartitionPlan& p, cudaStream_t stream) {
    k_binscan<<<nbins, PART_THREADS, 0, stream>>>(d_cta_cnt, (uint32_t)p.grid, nbins, d_bin_count, d_bin_start, d_done_counter);
    k_scatter<<<p.grid, PART_THREADS, p.smem, stream>>>(d_assign, d_hdr, nbins, p.nbits, d_cta_cnt, d_bin_start, d_order, d_zero_next);
    return cudaGetLastError();
}

}  // namespace sbr
