// synth_common.h -- the synthetic read generator, one definition for host and device.
//
// Record i of a stream is a pure function of (seed, mate, i): any slice can be produced anywhere
// (GPU shard, CPU sample for the baseline, a test) and is bit-identical.  Shapes follow
// BASELINE.json's synthetic configs: fixed-length reads, 12-digit headers,
// LF line ends, bare '+' separator, so every record is canonical.
//
//   FASTQ record:  "@S%012llu/M\n" + bases + "\n+\n" + quals + "\n"      (17 + L + 1 + 2 + L + 1 bytes)
//   FASTA record:  ">S%012llu\n"   + bases + "\n"                         (15 + L + 1 bytes)
//
//   h   = mix(seed ^ mix(2*i + mate));  r(j) = mix(h + j)
//   r(0): 90 % of reads are "tagged": they start with barcode row (r0>>8) % N carrying
//         e substitutions (70 % e=0, 15 % e=1, 10 % e=2, 5 % e=3) at distinct positions, each to a
//         different base; the other 10 % keep their random prefix.
//   r(1): substitution positions / replacement bases.       r(2): 0-2 'N' bases per read (~14 % of reads).
//   r(16 + p/32): two bits per base.                        r(64 + p/8): quality byte p, '!'..'I'.
#pragma once
#include <stdint.h>

#include "sbr_common.cuh"

namespace sbr {

struct SynthParams {
    uint64_t seed;
    uint32_t fmt;
    uint32_t read_len;
    uint32_t mate;
    uint32_t n_barcodes;
    uint32_t bc_len;
    uint32_t rec_bytes;
    const uint8_t* barcodes;  // same address space as the caller
};

__host__ __device__ __forceinline__ uint32_t synth_rec_bytes(uint32_t fmt, uint32_t read_len) {
    return fmt == SBR_FMT_FASTQ ? 17u + (read_len + 1u) + 2u + (read_len + 1u) : 15u + read_len + 1u;
}

// writes one record (p.rec_bytes bytes) to out
__host__ __device__ inline void synth_record(const SynthParams& p, uint64_t i, uint8_t* out) {
    const uint32_t L = p.read_len;
    const uint64_t h = mix64(p.seed ^ mix64(2ull * i + p.mate));
    uint32_t w = 0;
    out[w++] = p.fmt == SBR_FMT_FASTQ ? '@' : '>';
    out[w++] = 'S';
    {
        uint64_t v = i;
        for (int d = 11; d >= 0; --d) { out[w + d] = (uint8_t)('0' + (v % 10)); v /= 10; }
        w += 12;
    }
    if (p.fmt == SBR_FMT_FASTQ) { out[w++] = '/'; out[w++] = (uint8_t)('1' + p.mate); }
    out[w++] = '\n';
    uint8_t* seq = out + w;
    const char ACGT[4] = {'A', 'C', 'G', 'T'};
    uint64_t word = 0;
    for (uint32_t q = 0; q < L; ++q) {
        if ((q & 31u) == 0) word = mix64(h + 16 + (q >> 5));
        seq[q] = (uint8_t)ACGT[(word >> (2 * (q & 31u))) & 3u];
    }
    const uint64_t r0 = mix64(h + 0);
    const uint32_t k = p.bc_len < L ? p.bc_len : L;
    if (p.n_barcodes != 0 && (r0 % 10u) != 0u) {
        const uint8_t* bc = p.barcodes + (uint64_t)((r0 >> 8) % p.n_barcodes) * p.bc_len;
        for (uint32_t q = 0; q < k; ++q) seq[q] = bc[q];
        const uint32_t cls = (uint32_t)((r0 >> 40) % 100u);
        uint32_t e = cls < 70 ? 0u : (cls < 85 ? 1u : (cls < 95 ? 2u : 3u));
        if (e > k) e = k;
        const uint64_t r1 = mix64(h + 1);
        uint64_t used = 0;
        for (uint32_t s = 0; s < e; ++s) {
            uint32_t pick = (uint32_t)((r1 >> (8 * s)) & 0xFFu) % (k - s);
            uint32_t pos = 0;
            for (uint32_t q = 0; q < k; ++q) {  // pick-th unused position (k <= 64)
                if ((used >> q) & 1ull) continue;
                if (pick == 0) { pos = q; break; }
                --pick;
            }
            used |= 1ull << pos;
            uint32_t alt = (uint32_t)((r1 >> (32 + 2 * s)) & 3u) % 3u;
            uint8_t old = seq[pos];
            for (int c = 0; c < 4; ++c) {  // alt-th base different from the old one
                if ((uint8_t)ACGT[c] == old) continue;
                if (alt == 0) { seq[pos] = (uint8_t)ACGT[c]; break; }
                --alt;
            }
        }
    }
    {
        const uint64_t r2 = mix64(h + 2);
        const uint32_t sel = (uint32_t)(r2 & 0xFFu);
        const uint32_t nN = sel < 3 ? 2u : (sel < 36 ? 1u : 0u);
        if (nN >= 1 && L) seq[(uint32_t)((r2 >> 8) % L)] = 'N';
        if (nN >= 2 && L) seq[(uint32_t)((r2 >> 36) % L)] = 'N';
    }
    w += L;
    out[w++] = '\n';
    if (p.fmt == SBR_FMT_FASTQ) {
        out[w++] = '+';
        out[w++] = '\n';
        for (uint32_t q = 0; q < L; ++q) {
            if ((q & 7u) == 0) word = mix64(h + 64 + (q >> 3));
            out[w++] = (uint8_t)('!' + (uint32_t)((word >> (8 * (q & 7u))) & 0xFFu) % 41u);
        }
        out[w++] = '\n';
    }
}

}  // namespace sbr
