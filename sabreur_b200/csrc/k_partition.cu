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
//             its bins in shared memory while it writes the assignments -> cnt[bin][slice]
//   k_binscan one block per bin: exclusive scan of cnt[bin][*] in place, bin totals -> bin_count; the last
//             block to finish scans bin_count into bin_start
//   k_scatter same slices; per warp a running cursor per bin in shared memory; __match_any_sync
//             ranks equal bins inside each 32-record step so ties keep input order
//                                                                   (reads assign again, writes 4 B/record)
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

// One block per bin: in-place exclusive scan over the V slices of that bin (every thread owns a contiguous
// chunk of slices).  The last block to finish turns the bin totals into bin_start (exclusive scan over bins).
__global__ void __launch_bounds__(PART_THREADS) k_binscan(uint32_t* __restrict__ cnt, uint32_t V, uint32_t nbins,
                                                          uint32_t* __restrict__ bin_count,
                                                          uint32_t* __restrict__ bin_start /* nbins + 1 */,
                                                          uint32_t* __restrict__ done_counter) {
    __shared__ uint32_t s_warp[PART_WARPS];
    __shared__ uint32_t s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* row = cnt + (size_t)blockIdx.x * V;
    // every warp owns a contiguous run of slices and walks it 32 at a time (coalesced)
    const uint32_t per = (((V + PART_WARPS - 1) / PART_WARPS) + 31u) & ~31u;
    const uint32_t w0 = min((uint32_t)warp * per, V), w1 = min(w0 + per, V);
    uint32_t sum = 0;
    for (uint32_t i = w0 + lane; i < w1; i += 32) sum += row[i];
    sum = warp_sum(sum);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    uint32_t carry = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < PART_WARPS; ++w) {
        uint32_t t = s_warp[w];
        if (w < warp) carry += t;
        tot += t;
    }
    for (uint32_t i0 = w0; i0 < w1; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint32_t x = i < w1 ? row[i] : 0u;
        const uint32_t inc = warp_incl_scan(x, lane);
        if (i < w1) row[i] = carry + inc - x;
        carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (threadIdx.x == 0) {
        bin_count[blockIdx.x] = tot;
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

__global__ void __launch_bounds__(PART_THREADS) k_scatter(const uint16_t* __restrict__ assign,
                                                          const BatchHeader* __restrict__ hdr, uint32_t nbins,
                                                          const uint32_t* __restrict__ cnt /* scanned */,
                                                          const uint32_t* __restrict__ bin_start,
                                                          uint32_t* __restrict__ order) {
    extern __shared__ uint32_t s_cur[];  // [PART_WARPS][nbins] running cursor per bin
    const uint32_t n = hdr->n_records;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t V = gridDim.x * PART_WARPS, v = blockIdx.x * PART_WARPS + warp;
    uint32_t* cur = s_cur + (size_t)warp * nbins;
    uint32_t lo, hi;
    slice_range(n, V, v, lo, hi);
    if (lo >= hi) return;  // warp-uniform
    for (uint32_t b = lane; b < nbins; b += 32) cur[b] = bin_start[b] + cnt[(size_t)b * V + v];
    __syncwarp();
    const uint32_t lt = (1u << lane) - 1u;
    constexpr int U = 4;  // the loads of four 32-record groups are in flight together; ranking stays in order
    for (uint32_t g0 = lo; g0 < hi; g0 += 32 * U) {
        uint32_t bins[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = g0 + u * 32 + lane;
            bins[u] = i < hi ? (uint32_t)assign[i] : 0xFFFFFFFFu;  // dead lanes form their own group
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = g0 + u * 32 + lane;
            if (g0 + u * 32 >= hi) break;  // warp-uniform
            const bool live = i < hi;
            const uint32_t bin = bins[u];
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, bin);
            const uint32_t rank = __popc(peers & lt);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (live && lane == leader) {
                base = cur[bin];
                cur[bin] = base + __popc(peers);
            }
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (live) order[base + rank] = i;
            __syncwarp();
        }
    }
}

}  // namespace

struct PartitionPlan {
    int grid;      // blocks of k_hist / k_scatter
    uint32_t V;    // warp slices
    size_t smem;   // dynamic shared bytes
};

PartitionPlan partition_plan(uint32_t nbins, uint32_t max_records, int sms) {
    PartitionPlan p;
    size_t smem = (size_t)PART_WARPS * nbins * sizeof(uint32_t);
    size_t fit = (200 * 1024) / (smem ? smem : 1);
    int per_sm = fit < 1 ? 1 : (fit > 8 ? 8 : (int)fit);
    uint32_t want = (max_records + PART_THREADS * 4 - 1) / (PART_THREADS * 4);  // >= 4 steps of 32 per warp
    uint32_t cap_mem = ((2u << 20) / nbins) / PART_WARPS;  // keep cnt[nbins][V] around 8 MB
    uint32_t g = min(want, (uint32_t)(sms * per_sm));
    g = min(g, max(cap_mem, (uint32_t)sms));
    p.grid = (int)max(1u, g);
    p.V = (uint32_t)p.grid * PART_WARPS;
    p.smem = smem;
    return p;
}

cudaError_t partition_prepare() {
    return cudaFuncSetAttribute(k_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
}

cudaError_t partition_launch(const uint16_t* d_assign, const BatchHeader* d_hdr, uint32_t nbins, uint32_t* d_cnt,
                             uint32_t* d_bin_count, uint32_t* d_bin_start, uint32_t* d_order, uint32_t* d_done_counter,
                             const PartitionPlan& p, cudaStream_t stream) {
    k_binscan<<<nbins, PART_THREADS, 0, stream>>>(d_cnt, p.V, nbins, d_bin_count, d_bin_start, d_done_counter);
    k_scatter<<<p.grid, PART_THREADS, p.smem, stream>>>(d_assign, d_hdr, nbins, d_cnt, d_bin_start, d_order);
    return cudaGetLastError();
}

}  // namespace sbr
