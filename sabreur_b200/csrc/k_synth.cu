// k_synth.cu -- synthetic FASTQ/FASTA generator (device kernel + host twin) and barcode tables.
// Workload synthesis for BASELINE.json's synthetic configs; see synth_common.h for the definition.
#include <cstdio>
#include <cstring>
#include <vector>

#include "synth_common.h"

namespace sbr {

namespace {

constexpr int SYNTH_THREADS = 128;

// one thread builds one record in shared memory, the block then streams the staged records out with
// coalesced 16-byte stores (block base offsets are multiples of 128 * rec_bytes => 16-byte aligned)
__global__ void __launch_bounds__(SYNTH_THREADS) k_synth(SynthParams p, uint64_t first, uint64_t n, uint8_t* out) {
    extern __shared__ __align__(16) uint8_t s_rec[];
    const uint64_t blk_first = (uint64_t)blockIdx.x * SYNTH_THREADS;
    if (blk_first >= n) return;
    const uint64_t left = n - blk_first;
    const uint32_t here = left < (uint64_t)SYNTH_THREADS ? (uint32_t)left : (uint32_t)SYNTH_THREADS;
    if (threadIdx.x < here) synth_record(p, first + blk_first + threadIdx.x, s_rec + (size_t)threadIdx.x * p.rec_bytes);
    __syncthreads();
    const size_t bytes = (size_t)here * p.rec_bytes;
    uint8_t* dst = out + blk_first * p.rec_bytes;
    const size_t vec = bytes / 16;
    for (size_t i = threadIdx.x; i < vec; i += SYNTH_THREADS)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(s_rec)[i];
    for (size_t i = vec * 16 + threadIdx.x; i < bytes; i += SYNTH_THREADS) dst[i] = s_rec[i];
}

// fallback for records too large to stage: one thread per record, direct global writes
__global__ void k_synth_direct(SynthParams p, uint64_t first, uint64_t n, uint8_t* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) synth_record(p, first + i, out + i * p.rec_bytes);
}

}  // namespace
}  // namespace sbr

using namespace sbr;

extern "C" int sbr_synth_table(uint64_t seed, uint32_t n_barcodes, uint32_t bc_len, uint32_t min_dist, uint8_t* out) {
    if (!out || bc_len == 0 || bc_len > 32 || n_barcodes == 0) return SBR_E_INVALID_ARG;
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    uint32_t got = 0;
    for (uint64_t j = 0; got < n_barcodes; ++j) {
        if (j > 200000000ull) return SBR_E_UNSUPPORTED;
        uint64_t c = mix64(seed + j);
        uint8_t cand[32];
        for (uint32_t p = 0; p < bc_len; ++p) cand[p] = (uint8_t)ACGT[(c >> (2 * p)) & 3u];
        bool ok = true;
        for (uint32_t a = 0; a < got && ok; ++a) {
            uint32_t d = 0;
            for (uint32_t p = 0; p < bc_len; ++p) d += out[(size_t)a * bc_len + p] != cand[p];
            ok = d >= min_dist;
        }
        if (ok) { memcpy(out + (size_t)got * bc_len, cand, bc_len); ++got; }
    }
    return SBR_OK;
}

extern "C" uint32_t sbr_synth_record_bytes(const sbr_synth_spec* spec) {
    return spec ? synth_rec_bytes(spec->fmt, spec->read_len) : 0u;
}

static bool spec_ok(const sbr_synth_spec* s) {
    return s && (s->fmt == SBR_FMT_FASTA || s->fmt == SBR_FMT_FASTQ) && s->read_len > 0 && s->bc_len <= 64 &&
           (s->n_barcodes == 0 || s->barcodes) && s->mate <= 1;
}

extern "C" int sbr_synth_host(const sbr_synth_spec* spec, uint64_t first_record, uint64_t n_records, uint8_t* out) {
    if (!spec_ok(spec) || !out) return SBR_E_INVALID_ARG;
    SynthParams p{spec->seed, spec->fmt, spec->read_len, spec->mate, spec->n_barcodes, spec->bc_len,
                  synth_rec_bytes(spec->fmt, spec->read_len), spec->barcodes};
    for (uint64_t i = 0; i < n_records; ++i) synth_record(p, first_record + i, out + i * p.rec_bytes);
    return SBR_OK;
}

extern "C" int sbr_synth_device(int device, const sbr_synth_spec* spec, uint64_t first_record, uint64_t n_records,
                                uint8_t* d_out, void* stream) {
    if (!spec_ok(spec) || !d_out) return SBR_E_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return SBR_E_NO_DEVICE;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* d_tab = nullptr;
    size_t tab_bytes = (size_t)spec->n_barcodes * spec->bc_len;
    if (tab_bytes) {
        if (cudaMalloc(&d_tab, tab_bytes) != cudaSuccess) return SBR_E_NOMEM;
        if (cudaMemcpyAsync(d_tab, spec->barcodes, tab_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            cudaFree(d_tab);
            return SBR_E_CUDA;
        }
    }
    SynthParams p{spec->seed, spec->fmt, spec->read_len, spec->mate, spec->n_barcodes, spec->bc_len,
                  synth_rec_bytes(spec->fmt, spec->read_len), d_tab};
    cudaError_t e = cudaSuccess;
    size_t smem = (size_t)SYNTH_THREADS * p.rec_bytes;
    // launch in chunks so a grid never exceeds 2^31-1 blocks
    const uint64_t chunk = (uint64_t)SYNTH_THREADS * (1u << 20);
    for (uint64_t done = 0; done < n_records && e == cudaSuccess; done += chunk) {
        uint64_t n = n_records - done < chunk ? n_records - done : chunk;
        uint8_t* dst = d_out + done * p.rec_bytes;
        if (smem <= 200 * 1024) {
            e = cudaFuncSetAttribute(k_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) break;
            unsigned grid = (unsigned)((n + SYNTH_THREADS - 1) / SYNTH_THREADS);
            k_synth<<<grid, SYNTH_THREADS, smem, st>>>(p, first_record + done, n, dst);
        } else {
            unsigned grid = (unsigned)((n + 127) / 128);
            k_synth_direct<<<grid, 128, 0, st>>>(p, first_record + done, n, dst);
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (d_tab) cudaFree(d_tab);
    return e == cudaSuccess ? SBR_OK : SBR_E_CUDA;
}
