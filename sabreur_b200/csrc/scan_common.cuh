// scan_common.cuh -- pieces shared by the two record-scan kernels (k_scan_fast.cu: region-parallel,
// k_scan.cu: decoupled look-back).  Record rules restate needletail 0.5.1's observable behaviour as
// used by src/demux.rs:23,49,85-86,101,114 (see oracle/sabreur_oracle.c for the CPU restatement).
#pragma once
#include "sbr_common.cuh"

namespace sbr {

constexpr uint32_t NOPOS = 0xFFFFFFFFu;
constexpr int GHOSTS = 4;  // list[GHOSTS + i] = i-th newline of the tile; list[GHOSTS-1-g] = g-th newline before it

// one region of the fast scan = the contiguous tiles one CTA streams through
struct RegionInfo {
    uint32_t nl_own;         // newlines inside the region's own byte range
    uint32_t nrec;           // records owned by the region (FASTQ: start byte inside; FASTA: preceding '\n' inside)
    uint32_t phase;          // FASTQ: (global newline index of the region's first newline) mod 4 as detected; 0xFF = unknown
    uint32_t flags;          // bit 0: the region met something the fast path does not handle
    uint32_t last_start_nl;  // FASTA: newlines of the region up to and including the one before its last record start
    uint32_t last_start_off; // FASTA: offset of the region's last record start
    uint32_t pad[2];
};
static_assert(sizeof(RegionInfo) == 32, "RegionInfo is 32 bytes");
constexpr uint32_t PHASE_UNKNOWN = 0xFFu;
constexpr int MAX_REGIONS = 1024;

#ifdef __CUDACC__

// byte at batch position pos: from the staged tile when it is inside, else from global memory (L2)
struct ByteSrc {
    const uint8_t* sm;
    const uint8_t* text;
    uint32_t lo, hi;
    __device__ __forceinline__ uint32_t operator()(uint32_t pos) const {
        return (pos >= lo && pos < hi) ? sm[pos - lo] : text[pos];
    }
};

// 2-bit pack of k (<= 16) bases starting at s: codes | invalid << 32
__device__ __forceinline__ uint64_t pack_prefix(const ByteSrc& B, uint32_t s, uint32_t k) {
    uint32_t codes = 0, inval = 0;
    for (uint32_t q = 0; q < k; ++q) {
        uint32_t c = B(s + q);
        uint32_t code = (c >> 1) & 3u;
        uint32_t canon = (0x47544341u >> (8 * code)) & 0xFFu;  // code 0..3 -> 'A','C','T','G'
        codes |= code << (2 * q);
        inval |= (uint32_t)(c != canon) << q;
    }
    return (uint64_t)codes | ((uint64_t)inval << 32);
}

// One FASTQ record given its four newline positions.  Returns SBR_ST_* violation bits (never
// INVALID_START: the '@' checks belong to the callers) and the record's scan word.
__device__ __forceinline__ uint32_t fastq_eval(const ByteSrc& B, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3,
                                               uint32_t bc_len, uint32_t packed, uint64_t& key) {
    uint32_t bits = 0;
    const uint32_t s = p0 + 1;
    const uint32_t cr0 = (p0 > 0 && B(p0 - 1) == '\r') ? 1u : 0u;  // the header line is never empty ('@')
    const uint32_t cr1 = (p1 > s && B(p1 - 1) == '\r') ? 1u : 0u;
    const uint32_t cr2 = (p2 > p1 + 1 && B(p2 - 1) == '\r') ? 1u : 0u;
    const uint32_t cr3 = (p3 > p2 + 1 && B(p3 - 1) == '\r') ? 1u : 0u;
    const uint32_t seq_len = (p1 - cr1) - s;
    const uint32_t qual_len = (p3 - cr3) - (p2 + 1);
    if (!(p1 + 1 < p2 && B(p1 + 1) == '+')) bits |= SBR_ST_INVALID_SEP;
    if (seq_len != qual_len) bits |= SBR_ST_UNEQUAL_LEN;
    key = 0;
    if (seq_len < bc_len) {
        bits |= SBR_ST_SHORT_READ;
        key |= KEY_SHORT | (packed ? 0ull : (uint64_t)s);
    } else {
        key |= packed ? pack_prefix(B, s, bc_len) : (uint64_t)s;
    }
    const bool sep_bare = ((p2 - cr2) - (p1 + 1)) == 1u;
    if (cr0 | cr1 | cr2 | cr3 | (uint32_t)!sep_bare) key |= KEY_NONCANON;
    return bits;
}

// One FASTA record starting at `start` ('>'); hend_hint = position of the header's newline when known.
// Returns 0 when bc_len bases were found, 1 when the record ended first (short read), 2 when the batch
// ended first (short read only if the batch is the end of the stream).
__device__ __forceinline__ uint32_t fasta_eval(const ByteSrc& B, uint32_t nbytes, uint32_t start, uint32_t hend_hint,
                                           uint32_t k, uint32_t packed, uint64_t& key) {
    uint32_t h = hend_hint;
    if (h == NOPOS) {
        h = start;
        while (h < nbytes && B(h) != '\n') ++h;
    }
    uint32_t q = h + 1, got = 0, codes = 0, inval = 0, first_seq = NOPOS, why = 2;
    if (k && q < nbytes && B(q) == '>') {  // the next record starts right behind the header: an empty sequence
        key = KEY_SHORT | (packed ? 0ull : (uint64_t)q);
        return 1;
    }
    while (got < k) {
        if (q >= nbytes) break;
        uint32_t c = B(q);
        if (c == '\n') {
            if (q + 1 >= nbytes) break;
            if (B(q + 1) == '>') { why = 1; break; }  // the record ends here
            ++q;
            continue;
        }
        if (c == '\r' && q + 1 < nbytes && B(q + 1) == '\n') { ++q; continue; }
        if (c == '\r' && q + 1 >= nbytes) break;
        if (first_seq == NOPOS) first_seq = q;
        if (packed) {
            uint32_t code = (c >> 1) & 3u;
            uint32_t canon = (0x47544341u >> (8 * code)) & 0xFFu;
            codes |= code << (2 * got);
            inval |= (uint32_t)(c != canon) << got;
        }
        ++got;
        ++q;
    }
    if (got < k) {
        key = KEY_SHORT | (packed ? 0ull : (uint64_t)(first_seq == NOPOS ? h + 1 : first_seq));
        return why;
    }
    key = packed ? ((uint64_t)codes | ((uint64_t)inval << 32)) : (uint64_t)first_seq;
    return 0;
}

// 16-bit mask of '>' bytes in a 16-byte chunk
__device__ __forceinline__ uint32_t gt_flags(uint32_t w) {
    uint32_t a = (w ^ 0x3E3E3E3Eu) & 0x7F7F7F7Fu;
    uint32_t b = a + 0x7F7F7F7Fu;
    return ~(b | w) & 0x80808080u;
}
__device__ __forceinline__ uint32_t gt_mask16(uint4 v) {
    uint32_t n0 = (gt_flags(v.x) * 0x00204081u) >> 28;
    uint32_t n1 = (gt_flags(v.y) * 0x00204081u) >> 28;
    uint32_t n2 = (gt_flags(v.z) * 0x00204081u) >> 28;
    uint32_t n3 = (gt_flags(v.w) * 0x00204081u) >> 28;
    return n0 | (n1 << 4) | (n2 << 8) | (n3 << 12);
}

__device__ __forceinline__ void report_bad(BatchHeader* hdr, uint32_t record, uint32_t bits) {
    atomicOr(&hdr->status, bits);
    unsigned long long packed = ((unsigned long long)record << 8) | bits;
    atomicMax(&hdr->first_bad_inv, ~packed);
}

#endif  // __CUDACC__

}  // namespace sbr
