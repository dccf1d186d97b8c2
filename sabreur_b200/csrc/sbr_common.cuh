// sbr_common.cuh -- shared device/host definitions for the sm_100a demux kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sabreur_gpu.h"

namespace sbr {

// ---- per-record scan word ("key"), written by the scan kernel, read by the match kernel ---------
// packed mode (bc_len <= 16, all-ACGT table):
//   [31:0]  2-bit codes of the first bc_len bases, base i at bits [2i+1:2i]; code = (ascii >> 1) & 3
//   [47:32] bit i set = base i is not one of A,C,G,T (mismatches every barcode, src/utils.rs:120)
// generic mode:
//   [31:0]  batch-relative offset of the first sequence byte
// both:
//   [48] short read (len(seq) < bc_len)   [49] non-canonical record
constexpr uint64_t KEY_SHORT = 1ull << 48;
constexpr uint64_t KEY_NONCANON = 1ull << 49;

// ---- scan kernel geometry -----------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_BYTES_PER_THREAD = 48;  // 3 x 16 B: odd chunk stride => conflict-free LDS.128
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_BYTES_PER_THREAD;  // 12288 B
constexpr int SCAN_STAGES = 4;
constexpr int SCAN_LIST_CAP = 2048;  // newline positions held per window (typical tile: ~150)
constexpr uint32_t SCAN_MAX_BATCH = 1u << 30;  // descriptor count field is 30 bits

// device header of one batch; zeroed (memset) before the scan kernel
struct BatchHeader {
    uint32_t ticket;       // dynamic tile tickets
    uint32_t n_records;    // records kept (<= max_records)
    uint32_t n_units;      // FASTQ: newline count; FASTA: record-start count
    uint32_t status;       // OR of SBR_ST_*
    unsigned long long first_bad_inv;  // ~min((record << 8) | bits); 0 = none
    uint32_t fmt;
    uint32_t total_records;  // records found before the max_records cap
    uint32_t nl_total;       // FASTA: newlines in the batch
    uint32_t nl_last_start;  // FASTA: newlines up to and including the one before the last record start
    uint32_t nl_at_cap;      // FASTA: same, for record index max_records
    uint32_t mode;           // 1: per-region segments (fast scan), 0: dense arrays (look-back scan)
    uint32_t spec_fail;      // fast scan gave up: the look-back scan must run
    uint32_t consumed;       // fast scan: end of the last complete record
    uint32_t n_regions;
    uint32_t arrive;         // fast scan: CTAs that have finished their region (the last one runs the epilogue)
};
static_assert(sizeof(BatchHeader) == 64, "BatchHeader is one 64-byte line");

// packed barcode table in constant-size device memory (<= 4096 rows)
struct PackedTable {
    uint32_t n;        // rows
    uint32_t k;        // match window (row 0 length)
    uint32_t m;        // mismatches allowed
    uint32_t pos_mask; // even bits of positions < k
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

#ifdef __CUDACC__
// 0x80 in every byte of w that equals '\n' -- exact (no borrow artefacts), 3 ops
__device__ __forceinline__ uint32_t nl_flags(uint32_t w) {
    uint32_t a = (w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu;
    uint32_t b = a + 0x7F7F7F7Fu;          // bit 7 set iff low 7 bits of (byte ^ '\n') != 0
    return ~(b | w) & 0x80808080u;         // '\n' has bit 7 clear
}
// 16-bit newline mask of a 16-byte chunk, bit i = byte i
__device__ __forceinline__ uint32_t nl_mask16(uint4 v) {
    // (z * 0x00204081) >> 28 gathers bits 7,15,23,31 into a nibble in byte order
    uint32_t n0 = (nl_flags(v.x) * 0x00204081u) >> 28;
    uint32_t n1 = (nl_flags(v.y) * 0x00204081u) >> 28;
    uint32_t n2 = (nl_flags(v.z) * 0x00204081u) >> 28;
    uint32_t n3 = (nl_flags(v.w) * 0x00204081u) >> 28;
    return n0 | (n1 << 4) | (n2 << 8) | (n3 << 12);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// the same with an L2 eviction-priority hint: the text is read exactly once, so its lines are marked evict-first and
// leave the (much smaller) per-record arrays the scan writes resident in L2 for the match kernel
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// bulk prefetch global -> L2 (no shared memory involved): the tile after next is on its way while this one is scanned
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(policy) : "memory");
}

// stores that should stay in L2 until the next kernel has read them (the scan's per-record output, read by the match kernel)
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_keep_u32(uint32_t* p, uint32_t v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void st_keep_u64(uint64_t* p, uint64_t v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(policy) : "memory");
}

// single-copy-atomic 128-bit accesses for the look-back descriptors (SASS LDG/STG.E.128.STRONG.GPU)
__device__ __forceinline__ uint4 ld_desc(const uint4* p) {
    uint4 v;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%4];\n\tmov.b128 {%0, %1, %2, %3}, t;\n\t}"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(uint4* p, uint4 v) {
    asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2, %3, %4};\n\tst.relaxed.gpu.global.b128 [%0], t;\n\t}" ::"l"(p),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}
// streaming (read-once) loads / stores
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
#endif  // __CUDACC__

}  // namespace sbr
