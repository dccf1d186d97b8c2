// k_scan.cu -- kernel 1: FASTQ/FASTA record-boundary scan (+ fused prefix packing).
//
// Replaces the record discovery that needletail::parse_fastx_reader(...).next() performs one record
// at a time inside the reference's hot loops (src/demux.rs:23,49 / 85-86,101,114) and the slice
// `&record.seq()[..bc_len]` (src/demux.rs:54,102,115).
//
// One pass over the text, HBM-bound:
//   * persistent CTAs pull 12 KB tiles (dynamic tickets) through a 4-stage ring filled by 1-D bulk
//     async copies (cp.async.bulk -> SASS UBLKCP) completing on mbarriers;
//   * every thread turns 3 x 16 B of the staged tile (conflict-free LDS.128) into a 48-bit newline
//     mask with 3 logic ops + 1 multiply per word, popc + warp-shuffle scan gives the tile-local
//     line index of every newline, which are compacted into a shared list;
//   * the global line index comes from a single-pass decoupled look-back over 16-byte tile
//     descriptors {state|count, last three newline positions} read/written with single-copy-atomic
//     128-bit accesses, so records that straddle tiles are validated without re-reading text;
//   * one thread per record (the one whose 4th newline lies in the tile) validates the record the
//     way needletail does ('@', '+', equal lengths, CR handling), writes rec_off[r+1] and the packed
//     2-bit prefix key[r] that kernel 2 consumes as a coalesced 8-byte stream.
#include "scan_common.cuh"

namespace sbr {

namespace {

constexpr uint32_t INVALID_TILE = 0xFFFFFFFFu;
constexpr uint32_t DESC_AGG = 1u, DESC_INC = 2u;

struct ScanSmem {
    uint64_t full_bar[SCAN_STAGES];
    uint32_t stage_tile[SCAN_STAGES];
    uint32_t warp_tot[SCAN_THREADS / 32];
    uint32_t tail[3];       // last three newline positions of the tile, [0] most recent
    uint32_t base;          // newlines (FASTQ) / record starts (FASTA) before this tile
    uint32_t base_nl;       // FASTA: newlines before this tile
    uint32_t list[GHOSTS + SCAN_LIST_CAP];
};

struct ScanArgs {
    const uint8_t* text;
    uint32_t nbytes;
    uint32_t ntiles;
    uint32_t max_records;
    uint32_t bc_len;
    uint32_t packed;       // 1: key = packed prefix, 0: key = seq offset
    uint32_t final_batch;
    uint32_t gated;        // 1: run only if the fast scan gave up (hdr->spec_fail)
    BatchHeader* hdr;
    uint4* desc;
    uint32_t* rec_off;
    uint64_t* key;
};

// decoupled look-back over tile descriptors; executed by warp 0.
// in:  own aggregate (cnt, t0,t1,t2 = own last newline positions, most recent first)
// out: base = units before this tile, g[0..2] = the three most recent newline positions before
//      the tile (NOPOS when absent).  Publishes AGG first and INC at the end.
template <bool WITH_POS>
__device__ __forceinline__ uint32_t lookback(uint4* desc, uint32_t tile, uint32_t cnt, uint32_t t0, uint32_t t1,
                                             uint32_t t2, uint32_t g[3], int lane) {
    g[0] = g[1] = g[2] = NOPOS;
    uint32_t base = 0;
    if (tile != 0) {
        if (lane == 0) st_desc(desc + tile, make_uint4((DESC_AGG << 30) | cnt, t0, t1, t2));
        int have = 0;
        int look = (int)tile - 1;
        while (true) {
            int idx = look - lane;
            uint4 d = make_uint4(DESC_INC << 30, NOPOS, NOPOS, NOPOS);  // before tile 0: inclusive, empty
            if (idx >= 0) d = ld_desc(desc + idx);
            uint32_t first, need;
            while (true) {
                uint32_t st = d.x >> 30;
                uint32_t inc_mask = __ballot_sync(0xFFFFFFFFu, st == DESC_INC);
                uint32_t bad_mask = __ballot_sync(0xFFFFFFFFu, st == 0);
                first = inc_mask ? (uint32_t)(__ffs(inc_mask) - 1) : 32u;
                need = first < 32u ? ((2u << first) - 1u) : 0xFFFFFFFFu;
                if ((bad_mask & need) == 0) break;
                if (st == 0) d = ld_desc(desc + idx);  // idx >= 0 here (virtual descriptors are INC)
            }
            uint32_t c = ((need >> lane) & 1u) ? (d.x & 0x3FFFFFFFu) : 0u;
            base += warp_sum(c);
            if (WITH_POS) {
                uint32_t last = first < 32u ? first : 31u;
                for (uint32_t L = 0; L <= last && have < 3; ++L) {
                    uint32_t cL = __shfl_sync(0xFFFFFFFFu, d.x, L) & 0x3FFFFFFFu;
                    uint32_t a = __shfl_sync(0xFFFFFFFFu, d.y, L);
                    uint32_t b = __shfl_sync(0xFFFFFFFFu, d.z, L);
                    uint32_t e = __shfl_sync(0xFFFFFFFFu, d.w, L);
                    // an INC descriptor carries the last three newlines of the whole prefix
                    uint32_t avail = (a != NOPOS) + (b != NOPOS) + (e != NOPOS);
                    (void)cL;
                    if (avail >= 1 && have < 3) g[have++] = a;
                    if (avail >= 2 && have < 3) g[have++] = b;
                    if (avail >= 3 && have < 3) g[have++] = e;
                }
            }
            if (first < 32u) break;
            look -= 32;
        }
    }
    // inclusive descriptor: own newlines first, then what came before
    uint32_t i0 = t0, i1 = t1, i2 = t2;
    if (WITH_POS) {
        if (cnt == 0) { i0 = g[0]; i1 = g[1]; i2 = g[2]; }
        else if (cnt == 1) { i1 = g[0]; i2 = g[1]; }
        else if (cnt == 2) { i2 = g[0]; }
    }
    if (lane == 0) st_desc(desc + tile, make_uint4((DESC_INC << 30) | (base + cnt), i0, i1, i2));
    return base;
}

// FASTA flavour: descriptor = {state|record starts, newlines, -, -}; no positions needed.
__device__ __forceinline__ uint32_t lookback_fasta(uint4* desc, uint32_t tile, uint32_t cnt, uint32_t nl,
                                                   uint32_t* base_nl_out, int lane) {
    uint32_t base = 0, base_nl = 0;
    if (tile != 0) {
        if (lane == 0) st_desc(desc + tile, make_uint4((DESC_AGG << 30) | cnt, nl, 0u, 0u));
        int look = (int)tile - 1;
        while (true) {
            int idx = look - lane;
            uint4 d = make_uint4(DESC_INC << 30, 0u, 0u, 0u);
            if (idx >= 0) d = ld_desc(desc + idx);
            uint32_t first, need;
            while (true) {
                uint32_t st = d.x >> 30;
                uint32_t inc_mask = __ballot_sync(0xFFFFFFFFu, st == DESC_INC);
                uint32_t bad_mask = __ballot_sync(0xFFFFFFFFu, st == 0);
                first = inc_mask ? (uint32_t)(__ffs(inc_mask) - 1) : 32u;
                need = first < 32u ? ((2u << first) - 1u) : 0xFFFFFFFFu;
                if ((bad_mask & need) == 0) break;
                if (st == 0) d = ld_desc(desc + idx);
            }
            bool use = (need >> lane) & 1u;
            base += warp_sum(use ? (d.x & 0x3FFFFFFFu) : 0u);
            base_nl += warp_sum(use ? d.y : 0u);
            if (first < 32u) break;
            look -= 32;
        }
    }
    if (lane == 0) st_desc(desc + tile, make_uint4((DESC_INC << 30) | (base + cnt), base_nl + nl, 0u, 0u));
    *base_nl_out = base_nl;
    return base;
}

// ------------------------------------------------------------------------------------------------
// FASTQ: one thread per record whose final newline is list entry i (window-relative, may use ghosts)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fastq_record(const ScanArgs& A, const ByteSrc& B, const uint32_t* list, int i,
                                             uint32_t r) {
    const uint32_t p3 = list[GHOSTS + i], p2 = list[GHOSTS + i - 1], p1 = list[GHOSTS + i - 2],
                   p0 = list[GHOSTS + i - 3];
    uint64_t key;
    uint32_t bits = fastq_eval(B, p0, p1, p2, p3, A.bc_len, A.packed, key);
    // the start of this record is checked by the previous record's thread ("next start" below);
    // record 0 checks byte 0 itself
    if (r == 0 && B(0) != '@') bits |= SBR_ST_INVALID_START;
    if (r < A.max_records) {
        A.rec_off[r + 1] = p3 + 1;
        A.key[r] = key;
        if (key & KEY_NONCANON) atomicOr(&A.hdr->status, SBR_ST_NONCANONICAL);
    }
    if (bits) report_bad(A.hdr, r, bits);
    // the byte after a record must start the next one; attributed to record r+1 and ignored by the
    // host when r+1 is not a complete record of this batch (blank / partial tail)
    if (p3 + 1 < A.nbytes && B(p3 + 1) != '@') report_bad(A.hdr, r + 1, SBR_ST_INVALID_START);
}

// ------------------------------------------------------------------------------------------------
// FASTA: a record starts at byte 0 and after every '\n' that is followed by '>'.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fasta_record(const ScanArgs& A, const ByteSrc& B, uint32_t start, uint32_t hend_hint,
                                             uint32_t r) {
    uint64_t key;
    const uint32_t why = fasta_eval(B, A.nbytes, start, hend_hint, A.bc_len, A.packed, key);
    if (why == 1 || (why == 2 && A.final_batch)) report_bad(A.hdr, r, SBR_ST_SHORT_READ);
    if (r <= A.max_records) A.rec_off[r] = start;
    if (r < A.max_records) A.key[r] = key;
}

template <int FMT>
__global__ void __launch_bounds__(SCAN_THREADS, 3) k_scan(ScanArgs A) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    uint8_t* stage_buf = dyn_smem;  // SCAN_STAGES * SCAN_TILE
    ScanSmem& S = *reinterpret_cast<ScanSmem*>(dyn_smem + SCAN_STAGES * SCAN_TILE);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t G = gridDim.x;
    if (A.gated && A.hdr->spec_fail == 0) return;  // the fast scan's result stands

    auto issue = [&](int s, uint32_t tile) {  // thread 0 only
        uint32_t lo = tile * (uint32_t)SCAN_TILE;
        uint32_t valid = min((uint32_t)SCAN_TILE, A.nbytes - lo);
        uint32_t bytes = (valid + 15u) & ~15u;  // the text buffer is padded to 16 bytes
        mbar_expect_tx(&S.full_bar[s], bytes);
        bulk_g2s(stage_buf + s * SCAN_TILE, A.text + lo, bytes, &S.full_bar[s]);
    };

    if (tid == 0) {
        for (int s = 0; s < SCAN_STAGES; ++s) mbar_init(&S.full_bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < SCAN_STAGES; ++s) {  // static first round, tickets afterwards
            uint32_t t = (uint32_t)s * G + blockIdx.x;
            S.stage_tile[s] = t < A.ntiles ? t : INVALID_TILE;
            if (t < A.ntiles) issue(s, t);
        }
    }
    __syncthreads();

    for (uint32_t it = 0;; ++it) {
        const int s = it % SCAN_STAGES;
        const uint32_t parity = (it / SCAN_STAGES) & 1u;
        const uint32_t tile = S.stage_tile[s];
        if (tile == INVALID_TILE) break;
        uint32_t next_ticket = 0;
        if (tid == 0) next_ticket = atomicAdd(&A.hdr->ticket, 1u) + SCAN_STAGES * G;  // latency hidden below

        const uint32_t lo = tile * (uint32_t)SCAN_TILE;
        const uint32_t valid = min((uint32_t)SCAN_TILE, A.nbytes - lo);
        const uint8_t* sm = stage_buf + s * SCAN_TILE;
        mbar_wait(&S.full_bar[s], parity);

        // ---- newline mask of this thread's 48 bytes ---------------------------------------------
        const uint32_t toff = tid * SCAN_BYTES_PER_THREAD;
        const uint4* src = reinterpret_cast<const uint4*>(sm + toff);
        uint64_t m = (uint64_t)nl_mask16(src[0]) | ((uint64_t)nl_mask16(src[1]) << 16) |
                     ((uint64_t)nl_mask16(src[2]) << 32);
        uint64_t vmask = ~0ull;  // bytes of this thread's span that belong to the batch
        if (toff + SCAN_BYTES_PER_THREAD > valid) {
            uint32_t nv = valid > toff ? valid - toff : 0u;
            vmask = (1ull << nv) - 1ull;  // nv < 48
        }
        m &= vmask;
        // FASTA: record starts = newlines followed by '>' (bit 47 looks at the next thread's first byte)
        uint64_t R = 0;
        if (FMT == SBR_FMT_FASTA) {
            uint64_t gm = (uint64_t)gt_mask16(src[0]) | ((uint64_t)gt_mask16(src[1]) << 16) |
                          ((uint64_t)gt_mask16(src[2]) << 32);
            const uint32_t nxt = toff + SCAN_BYTES_PER_THREAD;  // tile offset of the byte after this thread's span
            uint32_t nb = 0;
            if (nxt < valid) nb = sm[nxt];
            else if (lo + nxt < A.nbytes && nxt == (uint32_t)SCAN_TILE) nb = A.text[lo + nxt];
            gm &= vmask;
            R = m & ((gm >> 1) | (nb == '>' ? (1ull << 47) : 0ull));
            // a newline that is the last byte of the batch starts nothing
        }
        const uint32_t c = __popcll(m), cr = __popcll(R);
        const uint32_t incl2 = warp_incl_scan(c | (cr << 16), lane);  // both counts fit 16 bits (tile = 12288 B)
        if (lane == 31) S.warp_tot[warp] = incl2;
        __syncthreads();  // (1)
        uint32_t excl2 = incl2 - (c | (cr << 16)), T2 = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) {
            uint32_t t = S.warp_tot[w];
            if (w < warp) excl2 += t;
            T2 += t;
        }
        const uint32_t excl = excl2 & 0xFFFFu, T = T2 & 0xFFFFu;
        const uint32_t excl_r = excl2 >> 16, TR = T2 >> 16;
        (void)excl_r; (void)TR; (void)cr;
        // ---- compact newline positions (window 0) + the tile's last three ------------------------
        {
            uint64_t mm = m;
            uint32_t idx = excl;
            while (mm) {
                uint32_t b = __ffsll((long long)mm) - 1;
                mm &= mm - 1;
                uint32_t pos = lo + toff + b;
                if (idx < (uint32_t)SCAN_LIST_CAP) S.list[GHOSTS + idx] = pos;
                if (idx + 3 >= T) S.tail[T - 1 - idx] = pos;
                ++idx;
            }
        }
        __syncthreads();  // (2)

        const ByteSrc B{sm, A.text, lo, lo + valid};

        if (FMT == SBR_FMT_FASTQ) {
            if (warp == 0) {
                uint32_t t0 = T >= 1 ? S.tail[0] : NOPOS, t1 = T >= 2 ? S.tail[1] : NOPOS,
                         t2 = T >= 3 ? S.tail[2] : NOPOS;
                uint32_t g[3];
                uint32_t base = lookback<true>(A.desc, tile, T, t0, t1, t2, g, lane);
                if (lane == 0) {
                    S.base = base;
                    S.list[GHOSTS - 1] = g[0];
                    S.list[GHOSTS - 2] = g[1];
                    S.list[GHOSTS - 3] = g[2];
                    if (tile == A.ntiles - 1) {  // whole batch counted
                        uint32_t total = (base + T) >> 2;
                        A.hdr->n_units = base + T;
                        A.hdr->total_records = total;
                        A.hdr->n_records = min(total, A.max_records);
                        A.hdr->fmt = SBR_FMT_FASTQ;
                        if (total > A.max_records) atomicOr(&A.hdr->status, SBR_ST_REC_OVERFLOW);
                    }
                    if (tile == 0) A.rec_off[0] = 0;
                }
            }
            __syncthreads();  // (3)
            const uint32_t base = S.base;
            // records finishing in this tile: entries i with (base + i) % 4 == 3
            const uint32_t i0 = (3u - base) & 3u;
            const uint32_t Tw = min(T, (uint32_t)SCAN_LIST_CAP);
            for (uint32_t i = i0 + 4u * tid; i < Tw; i += 4u * SCAN_THREADS)
                fastq_record(A, B, S.list, (int)i, (base + i) >> 2);
            // rare: more newlines than the list holds -> further windows, ghosts carried over
            for (uint32_t w0 = SCAN_LIST_CAP; w0 < T; w0 += SCAN_LIST_CAP) {
                __syncthreads();
                uint32_t keep = 0;
                if (tid < 3) keep = S.list[GHOSTS + SCAN_LIST_CAP - 1 - tid];
                __syncthreads();
                if (tid < 3) S.list[GHOSTS - 1 - tid] = keep;
                uint64_t mm = m;
                uint32_t idx = excl;
                while (mm) {
                    uint32_t b = __ffsll((long long)mm) - 1;
                    mm &= mm - 1;
                    if (idx >= w0 && idx < w0 + SCAN_LIST_CAP) S.list[GHOSTS + idx - w0] = lo + toff + b;
                    ++idx;
                }
                __syncthreads();
                const uint32_t wn = min(T - w0, (uint32_t)SCAN_LIST_CAP);
                const uint32_t j0 = (3u - (base + w0)) & 3u;
                for (uint32_t i = j0 + 4u * tid; i < wn; i += 4u * SCAN_THREADS)
                    fastq_record(A, B, S.list, (int)i, (base + w0 + i) >> 2);
            }
        } else {
            // ---- FASTA: record starts were ranked by the same scan (cR), so every owner of a
            // "\n>" newline knows its record index once the look-back has delivered the base ---------
            if (warp == 0) {
                uint32_t base_nl = 0;
                uint32_t base = lookback_fasta(A.desc, tile, TR, T, &base_nl, lane);
                if (lane == 0) {
                    S.base = base;
                    S.base_nl = base_nl;
                }
            }
            __syncthreads();  // (3)
            const uint32_t base = S.base, base_nl = S.base_nl;
            {
                uint64_t rr = R;
                uint32_t rank = excl_r;
                while (rr) {
                    uint32_t b = __ffsll((long long)rr) - 1;
                    rr &= rr - 1;
                    const uint32_t r = base + rank + 1u;  // record 0 starts at byte 0
                    const uint32_t below = __popcll(m & ((1ull << b) - 1ull));
                    const uint32_t idx = excl + below;  // tile-local index of this newline
                    uint64_t above = m & ~((2ull << b) - 1ull);
                    uint32_t hint = NOPOS;
                    if (above) hint = lo + toff + (__ffsll((long long)above) - 1);
                    else if (idx + 1 < min(T, (uint32_t)SCAN_LIST_CAP)) hint = S.list[GHOSTS + idx + 1];
                    fasta_record(A, B, lo + toff + b + 1, hint, r);
                    const uint32_t nl_upto = base_nl + idx + 1u;  // newlines up to and including this one
                    if (r == A.max_records) A.hdr->nl_at_cap = nl_upto;
                    if (rank + 1 == TR) atomicMax(&A.hdr->nl_last_start, nl_upto);
                    ++rank;
                }
                // CR before LF anywhere makes the batch non-canonical (conservative: includes the tail)
                uint64_t mm = m;
                bool crlf = false;
                while (mm) {
                    uint32_t b = __ffsll((long long)mm) - 1;
                    mm &= mm - 1;
                    uint32_t pos = lo + toff + b;
                    crlf |= pos > 0 && B(pos - 1) == '\r';
                }
                if (crlf) atomicOr(&A.hdr->status, SBR_ST_NONCANONICAL);
            }
            if (tid == 0) {
                atomicAdd(&A.hdr->nl_total, T);
                if (tile == 0) {
                    if (B(0) != '>') report_bad(A.hdr, 0, SBR_ST_INVALID_START);
                    fasta_record(A, B, 0, T > 0 ? S.list[GHOSTS] : NOPOS, 0);
                }
            }
        }

        __syncthreads();  // (4) all reads of the stage and of the list are done
        if (tid == 0) {
            uint32_t t = next_ticket < A.ntiles ? next_ticket : INVALID_TILE;
            S.stage_tile[s] = t;
            if (t != INVALID_TILE) issue(s, t);
        }
    }
}

// FASTA epilogue: record count from the last inclusive descriptor; batch-level canonicity
__global__ void k_fasta_finish(BatchHeader* hdr, const uint4* desc, uint32_t ntiles, uint32_t nbytes,
                               uint32_t max_records, uint32_t final_batch, uint32_t gated, uint32_t* rec_off) {
    if (gated && hdr->spec_fail == 0) return;
    uint32_t starts = (ntiles ? (desc[ntiles - 1].x & 0x3FFFFFFFu) : 0u) + 1u;  // + the record at byte 0
    uint32_t complete = final_batch ? starts : starts - 1u;
    uint32_t kept = min(complete, max_records);
    hdr->n_units = starts;
    hdr->total_records = complete;
    hdr->n_records = kept;
    hdr->fmt = SBR_FMT_FASTA;
    if (complete > max_records) atomicOr(&hdr->status, SBR_ST_REC_OVERFLOW);
    if (final_batch && kept == complete) rec_off[kept] = nbytes;  // else rec_off[kept] is the next record's start
    // canonical records hold exactly two newlines: header end and sequence end
    uint32_t nl_kept = kept < complete ? hdr->nl_at_cap : (final_batch ? hdr->nl_total : hdr->nl_last_start);
    if (nl_kept != 2u * kept) atomicOr(&hdr->status, SBR_ST_NONCANONICAL);
}

}  // namespace

size_t scan_smem_bytes() { return (size_t)SCAN_STAGES * SCAN_TILE + sizeof(ScanSmem); }

cudaError_t scan_prepare(int device, int* grid_out) {
    cudaError_t e;
    size_t smem = scan_smem_bytes();
    if ((e = cudaFuncSetAttribute(k_scan<SBR_FMT_FASTQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) !=
        cudaSuccess)
        return e;
    if ((e = cudaFuncSetAttribute(k_scan<SBR_FMT_FASTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) !=
        cudaSuccess)
        return e;
    int per_sm = 0, sms = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan<SBR_FMT_FASTQ>, SCAN_THREADS, smem)) !=
        cudaSuccess)
        return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    *grid_out = per_sm * sms;
    return cudaSuccess;
}

// Launches the scan on `stream`.  hdr must have been zeroed and desc[0..ntiles) cleared.
cudaError_t scan_launch(uint32_t fmt, const uint8_t* d_text, uint32_t nbytes, uint32_t max_records, uint32_t bc_len,
                        uint32_t packed, uint32_t final_batch, uint32_t gated, BatchHeader* hdr, uint4* desc,
                        uint32_t* rec_off, uint64_t* key, int grid_cap, cudaStream_t stream) {
    ScanArgs A;
    A.text = d_text;
    A.nbytes = nbytes;
    A.ntiles = (nbytes + SCAN_TILE - 1) / SCAN_TILE;
    A.max_records = max_records;
    A.bc_len = bc_len;
    A.packed = packed;
    A.final_batch = final_batch;
    A.gated = gated;
    A.hdr = hdr;
    A.desc = desc;
    A.rec_off = rec_off;
    A.key = key;
    if (A.ntiles == 0) return cudaSuccess;
    int grid = (int)min((uint32_t)grid_cap, A.ntiles);
    size_t smem = scan_smem_bytes();
    if (fmt == SBR_FMT_FASTQ) {
        k_scan<SBR_FMT_FASTQ><<<grid, SCAN_THREADS, smem, stream>>>(A);
    } else {
        k_scan<SBR_FMT_FASTA><<<grid, SCAN_THREADS, smem, stream>>>(A);
        k_fasta_finish<<<1, 1, 0, stream>>>(hdr, desc, A.ntiles, nbytes, max_records, final_batch, gated, rec_off);
    }
    return cudaGetLastError();
}

uint32_t scan_tiles(uint32_t nbytes) { return (nbytes + SCAN_TILE - 1) / SCAN_TILE; }

}  // namespace sbr
