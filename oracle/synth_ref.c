/*
 * synth_ref.c -- the synthetic workload generator of SURVEY.md 8(d), restated on the oracle's side.
 *
 * TEST INFRASTRUCTURE ONLY (see sabreur_oracle.h).  The reference arm of bench.py (--impl reference) builds its
 * input with this file, so that the process timing the CPU baseline never maps the product's CUDA library;
 * tests/test_oracle.py asserts that the bytes equal the product generator's (sbr_synth_host) bit for bit.
 *
 *   mix(x)  : splitmix64 finaliser
 *   table   : candidate j = mix(seed + j), base p = "ACGT"[(c >> 2p) & 3]; kept when its Hamming distance to every
 *             kept row is >= min_dist (greedy)
 *   record i: h = mix(seed ^ mix(2 i + mate)); r(j) = mix(h + j)
 *             bases   : two bits each from r(16 + p / 32)
 *             r(0)    : nine reads in ten start with barcode row (r0 >> 8) % N, carrying e substitutions
 *                       (70 % none, 15 % one, 10 % two, 5 % three) at distinct positions taken from r(1)
 *             r(2)    : zero to two 'N' bases
 *             quality : one byte each from r(64 + p / 8), '!' .. 'I'
 *   FASTQ "@S%012llu/M\n" bases "\n+\n" quals "\n";  FASTA ">S%012llu\n" bases "\n"
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static const char BASES[] = "ACGT";

int orc_synth_table(uint64_t seed, uint32_t n, uint32_t k, uint32_t min_dist, uint8_t* out) {
    uint32_t kept = 0;
    if (!out || !n || !k || k > 32) return -1;
    for (uint64_t j = 0; kept < n; ++j) {
        uint8_t row[32];
        uint64_t c = mix(seed + j);
        int far_enough = 1;
        if (j > 200000000ull) return -2;
        for (uint32_t p = 0; p < k; ++p) row[p] = (uint8_t)BASES[(c >> (2 * p)) & 3];
        for (uint32_t a = 0; a < kept && far_enough; ++a) {
            uint32_t d = 0;
            for (uint32_t p = 0; p < k; ++p) d += out[(size_t)a * k + p] != row[p];
            far_enough = d >= min_dist;
        }
        if (far_enough) memcpy(out + (size_t)kept++ * k, row, k);
    }
    return 0;
}

uint32_t orc_synth_record_bytes(int fastq, uint32_t L) { return fastq ? 17 + (L + 1) + 2 + (L + 1) : 15 + L + 1; }

static uint8_t* one_record(uint8_t* o, int fastq, uint32_t L, uint32_t mate, uint64_t seed, uint64_t i, const uint8_t* table,
                           uint32_t n_bc, uint32_t k) {
    const uint64_t h = mix(seed ^ mix(2 * i + mate));
    char head[32];
    int hl = fastq ? snprintf(head, sizeof head, "@S%012llu/%u\n", (unsigned long long)i, mate + 1)
                   : snprintf(head, sizeof head, ">S%012llu\n", (unsigned long long)i);
    memcpy(o, head, (size_t)hl);
    o += hl;
    uint8_t* s = o;
    for (uint32_t p = 0; p < L; p += 32) {
        uint64_t bits = mix(h + 16 + p / 32);
        for (uint32_t q = p; q < L && q < p + 32; ++q, bits >>= 2) s[q] = (uint8_t)BASES[bits & 3];
    }
    const uint64_t r0 = mix(h);
    const uint32_t win = k < L ? k : L;
    if (n_bc && r0 % 10 != 0) {
        memcpy(s, table + (size_t)((r0 >> 8) % n_bc) * k, win);
        const uint32_t cls = (uint32_t)((r0 >> 40) % 100);
        uint32_t e = cls < 70 ? 0 : cls < 85 ? 1 : cls < 95 ? 2 : 3;
        if (e > win) e = win;
        const uint64_t r1 = mix(h + 1);
        uint8_t taken[64] = {0};
        for (uint32_t t = 0; t < e; ++t) {
            uint32_t nth = (uint32_t)((r1 >> (8 * t)) & 0xFF) % (win - t), pos = 0;
            for (uint32_t q = 0; q < win; ++q) {  /* the nth position not substituted yet */
                if (taken[q]) continue;
                if (nth-- == 0) { pos = q; break; }
            }
            taken[pos] = 1;
            uint32_t alt = (uint32_t)((r1 >> (32 + 2 * t)) & 3) % 3;
            for (int b = 0; b < 4; ++b) {  /* the alt-th base that differs from the one in place */
                if ((uint8_t)BASES[b] == s[pos]) continue;
                if (alt-- == 0) { s[pos] = (uint8_t)BASES[b]; break; }
            }
        }
    }
    const uint64_t r2 = mix(h + 2);
    const uint32_t sel = (uint32_t)(r2 & 0xFF), n_N = sel < 3 ? 2 : sel < 36 ? 1 : 0;
    if (n_N >= 1 && L) s[(r2 >> 8) % L] = 'N';
    if (n_N >= 2 && L) s[(r2 >> 36) % L] = 'N';
    o += L;
    *o++ = '\n';
    if (fastq) {
        *o++ = '+';
        *o++ = '\n';
        for (uint32_t p = 0; p < L; p += 8) {
            uint64_t bits = mix(h + 64 + p / 8);
            for (uint32_t q = p; q < L && q < p + 8; ++q, bits >>= 8) *o++ = (uint8_t)('!' + (bits & 0xFF) % 41);
        }
        *o++ = '\n';
    }
    return o;
}

int orc_synth_reads(int fastq, uint32_t read_len, uint32_t mate, uint64_t seed, const uint8_t* table, uint32_t n_bc, uint32_t k,
                    uint64_t first, uint64_t n, uint8_t* out) {
    if (!out || !read_len || k > 64 || mate > 1 || (n_bc && !table)) return -1;
    for (uint64_t i = 0; i < n; ++i) out = one_record(out, fastq, read_len, mate, seed, first + i, table, n_bc, k);
    return 0;
}
