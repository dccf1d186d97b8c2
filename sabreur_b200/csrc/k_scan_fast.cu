// k_scan_fast.cu -- kernel 1 (fast path): region-parallel FASTQ/FASTA record-boundary scan.
//
// Replaces the record discovery of needletail::parse_fastx_reader(...).next() (src/demux.rs:23,49 /
// 85-86,101,114) and the slice &record.seq()[..bc_len] (src/demux.rs:54,102,115).
//
// Why not a global prefix scan: a single-pass decoupled look-back over 12 KB tiles resolves at most
// 32 tiles per L2 round trip, which caps a newline-count scan near 0.45 TB/s on B200 (measured, see
// DESIGN.md).  Here nothing on the critical path waits on another CTA:
//   * the batch is cut into one contiguous region per CTA; the CTA streams its region tile by tile with 1-D bulk
//     async copies (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier; the tile after the one being loaded
//     is pulled into L2 with a bulk prefetch, so that the single-buffered stage waits for an L2 hit only;
//   * inside a tile every WARP owns a contiguous 4.5 KB slice: it builds the newline masks of its slice,
//     compacts the newline positions into a warp-private list and evaluates the records that END in
//     its slice, two lanes per record.  The only things a warp needs from the others are how many
//     newlines precede its slice (its place in the 4-line cycle) and the last four newline positions
//     before it; both are published before the tile's barrier, so every warp carries the same load
//     between barriers;
//   * FASTQ: where in the 4-line cycle a region starts is not known locally ('@' and '+' are legal
//     quality characters), so the CTA DETECTS it from its first eight newlines (the one phase under
//     which the first whole record parses) and the batch's epilogue (scan_finish, run by the last CTA to
//     finish its region) VERIFIES every region's phase against the
//     exact prefix sum of the regions' newline counts.  A batch where any region is wrong, malformed,
//     short, over capacity or undecidable is flagged (hdr->spec_fail) and re-scanned by the exact
//     look-back kernel (k_scan.cu), which then also reports the precise violation.  Results are
//     therefore always those of the exact scan; only clean batches finish here.
//   * a record belongs to the region that holds its first byte; the CTA runs past the end of its region
//     to finish its last record, and skips the partial record at its start;
//   * per record: validate ('+' line, equal lengths, CR, next record's '@'), 2-bit pack the first bc_len
//     bases, write seg_rec_off / seg_key at the region-local index.  The match kernel compacts the
//     per-region segments into dense arrays while it reads the keys.
#include <cstdlib>

#include "scan_common.cuh"

namespace sbr {

namespace {

constexpr uint32_t UNKNOWN_L = 0xFFFFFFFFu;
#ifndef SBR_FAST_WARPS
#define SBR_FAST_WARPS 8
#endif
constexpr int FAST_WARPS = SBR_FAST_WARPS;  // warps (= slices of a tile) per CTA
constexpr int FAST_THREADS = FAST_WARPS * 32;
// A lane masks 48 bytes in each of FAST_UNITS units of its warp's slice; one warp scan serves all of them.
// Two geometries (SBR_SCAN_SINGLE selects at compile time):
//   double-buffered: 2 units per lane, 3 KB slices, 24 KB tiles, two stages (the next tile loads while this one is scanned)
//   single-buffered: 3 units per lane, 4.5 KB slices, 36 KB tiles, one stage (the per-slice bookkeeping is amortised over
//                    half as much again and a slice ends ~14 records, which fills the record lanes; the other
//                    resident CTAs cover a CTA's load)
#ifndef SBR_SCAN_SINGLE
#define SBR_SCAN_SINGLE 1
#endif
constexpr int FAST_UNITS = SBR_SCAN_SINGLE ? 3 : 2;
constexpr int UNIT_BYTES = 32 * SCAN_BYTES_PER_THREAD;  // 1536 B: one 48-byte chunk per lane
constexpr int WARP_SLICE = FAST_UNITS * UNIT_BYTES;     // 3072 / 4608 B
constexpr int FAST_TILE = FAST_WARPS * WARP_SLICE;      // 24 / 36 KB
constexpr int FAST_STAGES = SBR_SCAN_SINGLE ? 1 : 2;
// every stage is preceded by a copy of the last TAIL_PAD bytes of the previous tile, so that a record which began there is
// still addressable in shared memory at (stage base + position - tile start)
constexpr int TAIL_PAD = 512;
__host__ __device__ constexpr int stage_stride(int fmt) { return (fmt == SBR_FMT_FASTQ ? TAIL_PAD : 0) + FAST_TILE; }  // FASTA has no use for the pad
// Where does the time go?  Experimental builds cut the tile loop short (results are wrong: bench with SBR_DEBUG_ABLATE=1):
//   1: no record evaluation   2: no compaction either   3: loads, barriers and the next tile's issue only
#ifndef SBR_ABLATE
#define SBR_ABLATE 0
#endif
#ifndef SBR_PF_AHEAD
#define SBR_PF_AHEAD 1  // L2 prefetch distance in tiles (0 = none).  Measured (profiles/r2_experiments.md): none 215 us, 1: 204, 2 / 3 / 4: 214 / 231 / 264
#endif
constexpr int PF_AHEAD = SBR_PF_AHEAD;
#ifndef SBR_WL_CAP
#define SBR_WL_CAP (SBR_SCAN_SINGLE ? 284 : 188)
#endif
constexpr int WL_CAP = SBR_WL_CAP;
// FASTA keeps a second per-slice array (ridx): its lists are shorter (>= 24-byte lines on average, denser slices are the
// exact scan's job) so that four of its CTAs leave as much shared memory to the match kernel running beside them as
// four FASTQ CTAs do
#ifndef SBR_WL_CAP_FASTA
#define SBR_WL_CAP_FASTA (SBR_SCAN_SINGLE ? 192 : 128)
#endif
constexpr int WL_CAP_FASTA = SBR_WL_CAP_FASTA < WL_CAP ? SBR_WL_CAP_FASTA : WL_CAP;
__host__ __device__ constexpr int wl_cap(int fmt) { return fmt == SBR_FMT_FASTA ? WL_CAP_FASTA : WL_CAP; }
constexpr int CNT_BITS = FAST_UNITS == 2 ? 16 : 10;   // per-unit newline counts of a slice share one word (valid while the slice fits WL_CAP)
constexpr uint32_t CNT_MASK = (1u << CNT_BITS) - 1u;
constexpr int FASTA_CHUNKS = (WL_CAP_FASTA + 31) / 32;
constexpr uint32_t FASTA_START = 1u << 30;  // list entry flag: a FASTA record starts behind this newline
constexpr uint32_t POS_MASK = FASTA_START - 1u;  // the position part of a list entry (batches are at most 1 GiB)

template <int FMT>
struct FastSmem {
    uint64_t full_bar[FAST_STAGES];
    // per warp: [GHOSTS - 1 - g] = g-th most recent newline before the slice, [GHOSTS + i] = i-th newline of the slice
    uint32_t wl[FAST_WARPS][GHOSTS + wl_cap(FMT)];
    uint32_t wtot[2][FAST_WARPS];           // by tile parity: newlines | record starts << 16 of each warp slice
    uint32_t wtail[2][FAST_WARPS][GHOSTS];  // by tile parity: the slice's last four newline positions, [0] most recent
    uint32_t whead[2][FAST_WARPS][2];       // by tile parity: the slice's first two newline positions (FASTA looks ahead)
    uint16_t ridx[FMT == SBR_FMT_FASTA ? FAST_WARPS : 1][FMT == SBR_FMT_FASTA ? WL_CAP_FASTA + 4 : 2];  // FASTA: list index of the u-th record start
    uint32_t gh[2][GHOSTS];                 // by tile parity: the four most recent newlines before the tile
    uint32_t cnt[2];                        // newlines of the region before the tile
    uint32_t rcnt[2];                       // FASTA: record starts of the region before the tile
    uint32_t first[8];  // FASTQ: the region's first eight newline positions
    uint32_t nfirst;
    uint32_t h;            // FASTQ: phase of the region's first newline
    uint32_t synced;
    uint32_t first_owned;  // FASTQ: region-local index of the final newline of the first owned record
    uint32_t nl_own;       // newlines inside the region's own byte range
    uint32_t wmax[FAST_WARPS][2];  // FASTQ: per warp {records owned so far, end of the last one}
    uint32_t fail;
    uint32_t done;         // FASTQ: the record open at the region's end has been finished
    uint32_t crlf;
    uint32_t is_last;      // this CTA was the last of the grid to finish its region: it runs the batch's epilogue
    unsigned long long wlast[FAST_WARPS];  // FASTA, per warp: (newlines up to the one before its last record start) << 32 | offset
};

struct FastArgs {
    const uint8_t* text;
    uint32_t nbytes;
    uint32_t ntiles;
    uint32_t tiles_per_region;  // every region has this many tiles, the first `extra_regions` one more
    uint32_t extra_regions;
    uint32_t cap_r;  // record capacity of one region's segment
    uint32_t bc_len;
    uint32_t packed;
    uint32_t final_batch;
    uint32_t c7f, c0a, c80, c3e;  // byte-splat constants, passed in so that they live in registers (one LOP3 per step)
    uint32_t code_mask, inval_mask;  // the bc_len cut of the packed scan word
    BatchHeader* hdr;
    RegionInfo* info;
    uint32_t* seg_rec_off;
    uint64_t* seg_key;
    // the batch's epilogue (run by the last CTA to finish, see scan_finish)
    uint32_t max_records;
    uint32_t n_desc;
    uint32_t* region_base;    // [n_regions + 1]
    uint32_t* rec_off_dense;
    uint4* lookback_desc;
    uint32_t* stats;          // context-wide counters: [0] batches finished here, [1] batches handed to the exact scan
};

// 0x80 in every byte of w equal to the splatted byte `cx` (cx < 0x80): three instructions
__device__ __forceinline__ uint32_t eq_flags(uint32_t w, uint32_t c7f, uint32_t cx, uint32_t c80) {
    uint32_t a = (w & c7f) ^ cx;
    uint32_t b = a + c7f;
    return ~(b | w) & c80;
}
// the 48-bit mask of bytes equal to cx in this thread's three 16-byte chunks: lo = bytes 0..31, hi = bytes 32..47.
// dp4a gathers the four 0x80 flags of a word into a nibble in one instruction: 0x80 * {1,2,4,8} = nibble << 7.
__device__ __forceinline__ void eq_mask48(const uint4& v0, const uint4& v1, const uint4& v2, uint32_t c7f, uint32_t cx,
                                          uint32_t c80, uint32_t& lo, uint32_t& hi) {
    const uint32_t WL = 0x08040201u, WH = 0x80402010u;
    uint32_t x0 = __dp4a(eq_flags(v0.x, c7f, cx, c80), WL, __dp4a(eq_flags(v0.y, c7f, cx, c80), WH, 0u));
    uint32_t x1 = __dp4a(eq_flags(v0.z, c7f, cx, c80), WL, __dp4a(eq_flags(v0.w, c7f, cx, c80), WH, 0u));
    uint32_t x2 = __dp4a(eq_flags(v1.x, c7f, cx, c80), WL, __dp4a(eq_flags(v1.y, c7f, cx, c80), WH, 0u));
    uint32_t x3 = __dp4a(eq_flags(v1.z, c7f, cx, c80), WL, __dp4a(eq_flags(v1.w, c7f, cx, c80), WH, 0u));
    uint32_t x4 = __dp4a(eq_flags(v2.x, c7f, cx, c80), WL, __dp4a(eq_flags(v2.y, c7f, cx, c80), WH, 0u));
    uint32_t x5 = __dp4a(eq_flags(v2.z, c7f, cx, c80), WL, __dp4a(eq_flags(v2.w, c7f, cx, c80), WH, 0u));
    lo = (x0 >> 7) + (x1 << 1) + (x2 << 9) + (x3 << 17);  // disjoint bit fields: + == |
    hi = (x4 >> 7) + (x5 << 1);
}

// store through a shared-window address kept in a register (the compiler otherwise rebuilds the window base
// inside the compaction loops)
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// FASTQ phase detection from the first eight newlines of a region (thread 0).
// Hypothesis h: newline 0 is the (h+1)-th line end of its record.  Under h the first record that lies
// wholly behind a record end starts after newline s = (3 - h) & 3 and owns newlines s+1 .. s+4.
__device__ __forceinline__ bool detect_phase(const ByteSrc& B, const uint32_t* first, uint32_t nbytes, uint32_t& h_out) {
    int found = -1;
    for (uint32_t h = 0; h < 4; ++h) {
        uint32_t s = (3u - h) & 3u;
        uint32_t a = first[s], p0 = first[s + 1], p1 = first[s + 2], p2 = first[s + 3], p3 = first[s + 4];
        if (B(a + 1) != '@') continue;
        uint64_t key;
        uint32_t bits = fastq_eval(B, p0, p1, p2, p3, 0u, 0u, key);
        if (bits & (SBR_ST_INVALID_SEP | SBR_ST_UNEQUAL_LEN)) continue;
        if (p3 + 1 < nbytes && B(p3 + 1) != '@') continue;
        if (found < 0) found = (int)h;  // ties are settled by the epilogue's exact check (scan_finish)
    }
    if (found < 0) return false;
    h_out = (uint32_t)found;
    return true;
}

// 2-bit pack + validity of the first 4*KW bases at shared-window address `addr` (any alignment; the enclosing 4-byte
// words are read, up to 7 bytes past the window): codes | invalid << 32, cut to bc_len by the masks.
template <int KW>
__device__ __forceinline__ uint64_t pack_prefix_vec(uint32_t addr, uint32_t code_mask, uint32_t inval_mask) {
    const uint32_t sh = (addr & 3u) * 8u, wa = addr & ~3u;
    uint32_t codes = 0, inval = 0;
    uint32_t w = lds_u32(wa);
#pragma unroll
    for (int j = 0; j < KW; ++j) {
        const uint32_t wn = lds_u32(wa + 4u * (j + 1));
        const uint32_t x = __funnelshift_r(w, wn, sh);  // four bases, base 0 in the low byte
        w = wn;
        const uint32_t c = (x >> 1) & 0x03030303u;      // code = (ascii >> 1) & 3
        codes |= (((c * 0x00041041u) >> 18) & 0xFFu) << (8 * j);
        const uint32_t b0 = c & 0x01010101u, b1 = (c >> 1) & 0x01010101u;
        const uint32_t canon = 0x41414141u + 2u * b0 + 19u * b1 - 15u * (b0 & b1);  // 'A','C','T','G' from the code
        const uint32_t d = canon ^ x;
        const uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
        inval |= (__dp4a(nz, 0x08040201u, 0u) >> 7) << (4 * j);
    }
    return (uint64_t)(codes & code_mask) | ((uint64_t)(inval & inval_mask) << 32);
}

// a record that is not in canonical form: the general evaluation (rare, kept out of line).
// Returns the scan word with the violation bits in its top byte.
__device__ __noinline__ uint64_t fastq_slow(const uint8_t* sm, const uint8_t* text, uint32_t lo, uint32_t hi, uint32_t nbytes,
                                            uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t bc_len,
                                            uint32_t packed, uint32_t first_rec) {
    const ByteSrc B{sm, text, lo, hi};
    uint64_t key = 0;
    if (!(p0 < p1 && p1 < p2 && p2 < p3 && p3 < nbytes)) return (uint64_t)SBR_ST_INVALID_SEP << 56;  // positions from a failed slice
    uint32_t bits = fastq_eval(B, p0, p1, p2, p3, bc_len, packed, key);
    if (first_rec && B(0) != '@') bits |= SBR_ST_INVALID_START;
    if (p3 + 1 < nbytes && B(p3 + 1) != '@') bits |= SBR_ST_INVALID_START;
    return key | ((uint64_t)bits << 56);
}

// The batch's epilogue, run by the CTA that finishes its region last (an atomic ticket in the header, so no
// second launch): verifies every region against exact prefix sums and publishes the batch header.
//   * FASTQ: a region's detected phase must equal (newlines before the region) mod 4;
//   * any region that met something the fast path does not handle, or more records than max_records, hands the
//     whole batch to the exact look-back scan (hdr->spec_fail).
// Every thread owns RPT consecutive regions; `sc` is scratch shared memory (>= 32 words).
template <int FMT>
__device__ __forceinline__ void scan_finish(const FastArgs& A, uint32_t n_regions, uint32_t* sc) {
    constexpr int RPT = (MAX_REGIONS + FAST_THREADS - 1) / FAST_THREADS;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t nl[RPT], nr[RPT], ph[RPT], fl[RPT];
    uint32_t snl = 0, srec = 0;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
        const uint32_t c = tid * RPT + j;
        nl[j] = nr[j] = fl[j] = 0u;
        ph[j] = PHASE_UNKNOWN;
        if (c < n_regions) {  // written by other CTAs before their ticket: read past L1
            const uint4 v = __ldcg(reinterpret_cast<const uint4*>(A.info + c));
            nl[j] = v.x; nr[j] = v.y; ph[j] = v.z; fl[j] = v.w;
        }
        snl += nl[j];
        srec += nr[j];
    }
    if (tid == 0) { sc[16] = 0u; sc[17] = 0u; sc[18] = 0u; }  // fail, last region with records, newlines before it
    const uint32_t inl = warp_incl_scan(snl, lane), irec = warp_incl_scan(srec, lane);
    if (lane == 31) { sc[warp] = inl; sc[8 + warp] = irec; }
    __syncthreads();
    uint32_t nl_before = inl - snl, rec_before = irec - srec, total_nl = 0, total_rec = 0;
#pragma unroll
    for (int w = 0; w < FAST_WARPS; ++w) {
        const uint32_t a = sc[w], b = sc[8 + w];
        if (w < (int)warp) { nl_before += a; rec_before += b; }
        total_nl += a;
        total_rec += b;
    }
    bool bad = false;
    uint32_t my_last = 0, my_last_nlb = 0;
    bool have = false;
    {
        uint32_t nlb = nl_before, rb = rec_before;
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const uint32_t c = tid * RPT + j;
            if (c < n_regions) {
                if (fl[j] & 1u) bad = true;
                if (FMT == SBR_FMT_FASTQ) {
                    if (ph[j] == PHASE_UNKNOWN) bad |= nl[j] != 0;  // lines it could not place
                    else bad |= ph[j] != (nlb & 3u);                // detected phase vs the exact one
                }
                if (nr[j]) { have = true; my_last = c; my_last_nlb = nlb; }
                A.region_base[c] = rb;
            }
            nlb += nl[j];
            rb += nr[j];
        }
    }
    if (bad) sc[16] = 1u;
    if (have) atomicMax(&sc[17], my_last);
    __syncthreads();
    if (have && my_last == sc[17]) sc[18] = my_last_nlb;
    __syncthreads();
    BatchHeader* hdr = A.hdr;
    uint32_t consumed = __ldcg(&hdr->consumed);
    uint32_t nl_kept = total_nl;
    if (FMT == SBR_FMT_FASTA) {
        if (A.final_batch) consumed = A.nbytes;
        else {  // the last record start of a non-final batch opens an incomplete record
            total_rec -= 1;  // there is always record 0
            const uint4 L = __ldcg(reinterpret_cast<const uint4*>(A.info + sc[17]) + 1);  // {last_start_nl, last_start_off, -, -}
            consumed = L.x ? L.y : 0u;  // record 0 itself when no other start exists
            nl_kept = sc[18] + L.x;
        }
    }
    const bool fail = sc[16] != 0u || total_rec > A.max_records;
    if (fail)  // rare: the exact scan's tile descriptors must start out zeroed (the launch path does not memset them)
        for (uint32_t i = tid; i < A.n_desc; i += FAST_THREADS) A.lookback_desc[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        A.region_base[n_regions] = total_rec;
        if (A.stats) atomicAdd(&A.stats[fail ? 1 : 0], 1u);
        if (fail) {  // hand the batch to the exact scan with a clean header
            hdr->spec_fail = 1;
            hdr->mode = 0;
            hdr->status = 0;
            hdr->first_bad_inv = 0;
            hdr->n_records = 0;
            hdr->consumed = 0;
        } else {
            hdr->mode = 1;
            hdr->n_regions = n_regions;
            hdr->n_records = total_rec;
            hdr->total_records = total_rec;
            hdr->n_units = FMT == SBR_FMT_FASTQ ? total_nl : total_rec + (A.final_batch ? 0u : 1u);
            hdr->fmt = FMT;
            hdr->consumed = consumed;
            hdr->nl_total = total_nl;
            A.rec_off_dense[total_rec] = consumed;
            if (FMT == SBR_FMT_FASTA && nl_kept != 2u * total_rec) atomicOr(&hdr->status, SBR_ST_NONCANONICAL);
        }
    }
}

// KW: 32-bit words of text covering the match window (ceil(bc_len / 4), 1..4) in packed mode; 0 = unpacked keys.
// Register cap: 48 per thread costs no spills and no extra instructions over the 60 ptxas takes when left alone, and
// leaves a quarter of the register file to the previous batch's match + partition kernel, which runs beside the scan
// (4 scan CTAs + 1 match CTA per SM, see ctx.cu).
#ifndef SBR_SCAN_NREG
#define SBR_SCAN_NREG 48
#endif
template <int FMT, int KW>
__global__ void __maxnreg__(SBR_SCAN_NREG) k_scan_fast(FastArgs A) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    constexpr int STAGE_STRIDE = stage_stride(FMT);
    uint8_t* stage_buf = dyn_smem + (STAGE_STRIDE - FAST_TILE);  // stage s = stage_buf + s * STAGE_STRIDE, its tail pad right before it
    FastSmem<FMT>& S = *reinterpret_cast<FastSmem<FMT>*>(dyn_smem + FAST_STAGES * STAGE_STRIDE);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t c = blockIdx.x;
    // A.ntiles / tiles_per_region count FAST_TILE-sized tiles
    const uint32_t t_lo = c * A.tiles_per_region + min(c, A.extra_regions);
    const uint32_t t_hi = min(t_lo + A.tiles_per_region + (c < A.extra_regions ? 1u : 0u), A.ntiles);
    const bool idle = t_lo >= A.ntiles;  // (scan_fast_regions() hands every CTA at least one tile; kept for safety)
    const uint32_t region_lo = t_lo * (uint32_t)FAST_TILE;
    const uint32_t region_hi = min(t_hi * (uint32_t)FAST_TILE, A.nbytes);
    const uint32_t c7f = A.c7f, c0a = A.c0a, c80 = A.c80;
    uint32_t* const wl = S.wl[warp];
    const uint32_t wl_addr = smem_u32(wl + GHOSTS);  // shared-window address of the slice's entry 0

    const uint64_t text_policy = l2_evict_first_policy();
    const uint64_t keep_policy = l2_evict_last_policy();  // the per-record output: the match kernel reads it next
    auto issue = [&](int s, uint32_t tile) {  // thread 0 only
        uint32_t lo = tile * (uint32_t)FAST_TILE;
        uint32_t valid = min((uint32_t)FAST_TILE, A.nbytes - lo);
        uint32_t bytes = (valid + 15u) & ~15u;  // the text buffer is padded to 16 bytes
        mbar_expect_tx(&S.full_bar[s], bytes);
        bulk_g2s_hint(stage_buf + s * STAGE_STRIDE, A.text + lo, bytes, &S.full_bar[s], text_policy);
        if (FAST_STAGES == 1) {
            // single-buffered: the load of the NEXT tile can only start when this one has been scanned, so at least
            // bring the tile(s) behind it into L2 now.  (ncu: half of the text's sectors meet their own prefetch still in
            // flight when the load is issued one tile period later -- yet prefetching further ahead is slower: the tiles
            // needed next then queue behind less urgent ones.)  Nothing is prefetched past the tile a region runs over
            // into (t_hi).
            const uint32_t first = tile == t_lo ? 1u : (uint32_t)PF_AHEAD;  // a region's first load brings in all of them
            if (PF_AHEAD == 0) return;
            for (uint32_t k = first; k <= (uint32_t)PF_AHEAD; ++k) {
                const uint32_t nt = tile + k;
                if (nt > t_hi || nt >= A.ntiles) break;
                const uint32_t nlo = nt * (uint32_t)FAST_TILE;
                bulk_prefetch_l2(A.text + nlo, (min((uint32_t)FAST_TILE, A.nbytes - nlo) + 15u) & ~15u, text_policy);
            }
        }
    };

    if (tid == 0) {
        for (int s = 0; s < FAST_STAGES; ++s) mbar_init(&S.full_bar[s], 1);
        mbar_fence_init();
        S.nfirst = 0;
        S.h = 0;
        S.synced = (c == 0 || FMT == SBR_FMT_FASTA) ? 1u : 0u;  // a batch starts on a record boundary
        S.first_owned = (c == 0) ? 3u : UNKNOWN_L;
        S.cnt[0] = S.cnt[1] = 0;
        S.rcnt[0] = S.rcnt[1] = 0;
        S.nl_own = 0;
        for (int w = 0; w < FAST_WARPS; ++w) S.wmax[w][0] = S.wmax[w][1] = 0;
        S.fail = 0;
        S.done = 0;
        for (int w = 0; w < FAST_WARPS; ++w) S.wlast[w] = 0;
        S.crlf = 0;
        // "the newline before the region": a record whose final newline is the region's fourth starts at region_lo
        S.gh[0][0] = region_lo - 1u;  // == NOPOS for region 0
        for (int g = 1; g < GHOSTS; ++g) S.gh[0][g] = NOPOS;
        // record 0 of the batch has no predecessor to check its first byte
        if (FMT == SBR_FMT_FASTQ && c == 0 && A.text[0] != '@') S.fail = 1;
    }
    uint32_t* const seg_off_c = A.seg_rec_off + (size_t)c * A.cap_r;
    uint64_t* const seg_key_c = A.seg_key + (size_t)c * A.cap_r;
    __syncthreads();
    if (tid == 0 && !idle)
        for (int s = 0; s < FAST_STAGES; ++s)
            if (t_lo + s < t_hi) issue(s, t_lo + s);

    // Tile loop, ONE block barrier per tile while it is scanned (a second one at its end when single-buffered).
    // Before it every warp has compacted the newlines of its own
    // slice and published their number and its last four positions; after it every warp knows its place in the
    // line cycle, fetches the four newlines before its slice and evaluates the records ending in the slice.
    // The barrier of tile t+1 doubles as "everybody is done with tile t", after which thread 0 refills tile
    // t's stage.  Past the region's end (FASTQ only: the record left open there) the same body runs with one
    // more barrier and on-demand loads.
    for (uint32_t it = 0;; ++it) {
        const uint32_t tile = t_lo + it;
        if (idle || tile >= A.ntiles) break;
        const bool overrun = tile >= t_hi;
        if (overrun) {
            if (FMT == SBR_FMT_FASTA) break;
            if (FAST_STAGES > 1) __syncthreads();  // records of the previous tile are complete: S.done / S.synced are final
            if (S.done || !S.synced || S.fail) break;
        }
        const int s = it % FAST_STAGES;
        const uint32_t par = it & 1u;
        const uint32_t parity = (it / FAST_STAGES) & 1u;
        if (overrun && tid == 0) issue(s, tile);  // on demand: one extra tile per region, rarely more
        const uint32_t lo = tile * (uint32_t)FAST_TILE;
        const uint32_t valid = min((uint32_t)FAST_TILE, A.nbytes - lo);
        const uint8_t* sm = stage_buf + s * STAGE_STRIDE;
        mbar_wait(&S.full_bar[s], parity);
#if SBR_ABLATE >= 3
        if (true) {
            if (tid == 0 && tile + 1 >= t_hi) S.done = 1;  // (before the barrier: every thread must take the same exit)
            __syncthreads();
            if (warp == 0 && lane == 0 && tile + 1 < t_hi) issue(0, tile + 1);
            continue;
        }
#endif

        // ---- newline masks of this lane's 48 bytes in each unit of the warp's slice ---------------------
        const uint32_t woff = (uint32_t)warp * WARP_SLICE + (uint32_t)lane * SCAN_BYTES_PER_THREAD;
        uint32_t m_lo[FAST_UNITS], m_hi[FAST_UNITS];
        uint32_t cn2 = 0, c_all = 0;  // per-unit counts, CNT_BITS each; their sum
#pragma unroll
        for (int j = 0; j < FAST_UNITS; ++j) {
            const uint32_t uoff = woff + j * UNIT_BYTES;
            const uint4* src = reinterpret_cast<const uint4*>(sm + uoff);
            const uint4 v0 = src[0], v1 = src[1], v2 = src[2];
            eq_mask48(v0, v1, v2, c7f, c0a, c80, m_lo[j], m_hi[j]);
            if (uoff + SCAN_BYTES_PER_THREAD > valid) {
                uint32_t nv = valid > uoff ? valid - uoff : 0u;
                m_lo[j] &= nv >= 32 ? 0xFFFFFFFFu : ((1u << nv) - 1u);
                m_hi[j] &= nv > 32 ? ((1u << (nv - 32)) - 1u) : 0u;
            }
            const uint32_t cj = (uint32_t)(__popc(m_lo[j]) + __popc(m_hi[j]));
            cn2 |= cj << (CNT_BITS * j);
            c_all += cj;
        }
        const uint32_t n_w = __reduce_add_sync(0xFFFFFFFFu, c_all);  // newlines of the whole slice (exact even if fields overflow)
        const uint32_t incl_n = warp_incl_scan(cn2, lane);
        const uint32_t tot_n = __shfl_sync(0xFFFFFFFFu, incl_n, 31);
        const uint32_t ex_n = incl_n - cn2;
        uint32_t nr_w = 0;  // FASTA: record starts of the slice
        // read before the barrier: thread 0 may flip it right after that barrier, and every thread has to take
        // the same decision about the extra barrier below
        const bool was_synced = S.synced != 0;
        const bool fits = n_w <= (uint32_t)wl_cap(FMT);  // warp-uniform

        // ---- compact the slice's newline positions (ascending) into the warp's list ---------------------
        uint32_t start_bal[FASTA_CHUNKS] = {};  // FASTA: per 32 list entries, which ones are record starts
        if (fits && SBR_ABLATE < 2) {
            uint32_t unit_base = 0;  // newlines in the slice's earlier units
#pragma unroll
            for (int j = 0; j < FAST_UNITS; ++j) {
                const uint32_t idx = unit_base + ((ex_n >> (CNT_BITS * j)) & CNT_MASK);
                unit_base += (tot_n >> (CNT_BITS * j)) & CNT_MASK;
                uint32_t a = wl_addr + 4u * idx;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t mm = half ? m_hi[j] : m_lo[j];
                    const uint32_t pos0 = lo + woff + j * UNIT_BYTES + 32 * half;
                    while (mm) {
                        const uint32_t b = __ffs(mm) - 1;
                        mm &= mm - 1;
                        sts_u32(a, pos0 + b);
                        a += 4u;
                    }
                }
            }
            if (FMT == SBR_FMT_FASTA) {
                // one lane per list entry: the byte behind the newline ('>' = a record starts there: the entry gets
                // FASTA_START, bit 30 -- positions are below 2^30) and the byte before it (CR makes the batch non-canonical)
                __syncwarp();
                const uint32_t sm_a = smem_u32(sm);
                bool crlf = false;
#pragma unroll
                for (int ch = 0; ch < FASTA_CHUNKS; ++ch) {
                    if (ch * 32 >= (int)n_w) break;  // warp-uniform
                    const uint32_t jx = ch * 32 + lane;
                    bool st = false;
                    if (jx < n_w) {
                        const uint32_t pos = wl[GHOSTS + jx], o = pos - lo;
                        uint32_t nxt = 0, prv = 0;
                        if (o + 1 < valid) nxt = lds_u8(sm_a + o + 1);
                        else if (pos + 1 < A.nbytes) nxt = A.text[pos + 1];
                        if (o > 0) prv = lds_u8(sm_a + o - 1);
                        else if (pos > 0) prv = A.text[pos - 1];
                        crlf |= prv == '\r';
                        st = nxt == '>';
                        if (st) wl[GHOSTS + jx] = pos | FASTA_START;
                    }
                    start_bal[ch] = __ballot_sync(0xFFFFFFFFu, st);
                    if (st) S.ridx[warp][nr_w + __popc(start_bal[ch] & ((1u << lane) - 1u))] = (uint16_t)jx;  // rank -> list index
                    nr_w += __popc(start_bal[ch]);
                }
                if (crlf) S.crlf = 1;
            }
        } else if (lane == 0) {
            S.fail = 1;  // newline-dense slice: the exact scan's job
        }
        __syncwarp();
        // ---- publish: how many, and the last four ---------------------------------------------------------
        if (lane < GHOSTS) {
            const uint32_t nn = fits ? n_w : 0u;
            S.wtail[par][warp][lane] = nn > (uint32_t)lane ? wl[GHOSTS + nn - 1 - lane] : NOPOS;
            if (lane == 0) S.wtot[par][warp] = n_w | (nr_w << 16);
            if (FMT == SBR_FMT_FASTA && lane < 2) S.whead[par][warp][lane] = nn > (uint32_t)lane ? (wl[GHOSTS + lane] & POS_MASK) : NOPOS;
        }
        __syncthreads();  // the tile's barrier
        if (FAST_STAGES > 1 && tid == 0 && it >= 1 && !overrun) {  // tile it-1 is finished by everybody: refill its stage
            const uint32_t nt = tile - 1 + FAST_STAGES;
            if (nt < t_hi) issue((it - 1) % FAST_STAGES, nt);
        }
        if (FAST_STAGES > 1 && FMT == SBR_FMT_FASTQ && warp == FAST_WARPS - 2) {  // the tile's last TAIL_PAD bytes -> the pad in front of the other stage (the next tile)
            const uint4 v = *reinterpret_cast<const uint4*>(sm + FAST_TILE - TAIL_PAD + 16 * lane);
            *reinterpret_cast<uint4*>(stage_buf + (s ^ 1) * STAGE_STRIDE - TAIL_PAD + 16 * lane) = v;
        }
        // prefix over the warp slices by shuffles: lane w holds slice w's totals (newlines | starts << 16; a tile
        // has at most 24576 of either)
        uint32_t pre_n, T, pre_r = 0, TR = 0;
        {
            const uint32_t t = lane < FAST_WARPS ? S.wtot[par][lane] : 0u;
            uint32_t inc = t;
#pragma unroll
            for (int d = 1; d < FAST_WARPS; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += o;
            }
            const uint32_t pre = __shfl_sync(0xFFFFFFFFu, inc - t, warp);
            const uint32_t tot = __shfl_sync(0xFFFFFFFFu, inc, FAST_WARPS - 1);
            pre_n = pre & 0xFFFFu;
            T = tot & 0xFFFFu;
            if (FMT == SBR_FMT_FASTA) { pre_r = pre >> 16; TR = tot >> 16; }
        }
        const uint32_t base = S.cnt[par];
        // ---- the four newlines before the slice (lanes 0..3); warp 7 also those before the NEXT tile (4..7) --
        if (lane < GHOSTS || (warp == FAST_WARPS - 1 && lane < 2 * GHOSTS)) {
            const bool next_tile = lane >= GHOSTS;
            uint32_t need = (uint32_t)lane & (GHOSTS - 1);
            int v = (next_tile ? FAST_WARPS : warp) - 1;
            uint32_t val;
            while (true) {
                if (v < 0) { val = S.gh[par][need]; break; }
                const uint32_t cv = S.wtot[par][v] & 0xFFFFu;
                if (need < cv) { val = S.wtail[par][v][need]; break; }
                need -= cv;
                --v;
            }
            if (!next_tile) wl[GHOSTS - 1 - lane] = val;
            else {
                S.gh[par ^ 1u][lane - GHOSTS] = val;
                if (lane == GHOSTS) {
                    S.cnt[par ^ 1u] = base + T;
                    if (FMT == SBR_FMT_FASTA) S.rcnt[par ^ 1u] = S.rcnt[par] + TR;
                    if (!overrun) S.nl_own = base + T;
                }
            }
        }
        __syncwarp();

        if (FMT == SBR_FMT_FASTQ) {
            if (!was_synced) {  // uniform: only the first tile(s) of a region, until eight newlines were seen
                if (SBR_ABLATE < 2 && tid == 0 && !S.fail) {  // (a slice over capacity has no list to read from)
                    const ByteSrc B{sm, A.text, lo, lo + valid};
                    uint32_t nf = S.nfirst;
                    for (int v = 0; v < FAST_WARPS && nf < 8; ++v) {
                        const uint32_t nv = min(S.wtot[par][v] & 0xFFFFu, (uint32_t)wl_cap(FMT));
                        for (uint32_t i = 0; i < nv && nf < 8; ++i) S.first[nf++] = S.wl[v][GHOSTS + i];
                    }
                    S.nfirst = nf;
                    if (nf >= 8) {
                        uint32_t h;
                        if (detect_phase(B, S.first, A.nbytes, h)) {
                            S.h = h;
                            // the record ending at the first record-final newline is ours only if it starts
                            // exactly at the region's first byte
                            const uint32_t sfin = (3u - h) & 3u;
                            const bool own0 = (h == 0) && A.text[region_lo - 1] == '\n';
                            const uint32_t fo = own0 ? 3u : sfin + 4u;
                            S.first_owned = fo;
                            // records that ended in earlier tiles of this region (lines longer than a tile)
                            for (uint32_t L = fo; L < base; L += 4) {
                                uint64_t key;
                                const uint32_t p0 = S.first[L - 3], p1 = S.first[L - 2], p2 = S.first[L - 1], p3 = S.first[L];
                                uint32_t bits = fastq_eval(B, p0, p1, p2, p3, A.bc_len, A.packed, key);
                                if (p3 + 1 < A.nbytes && B(p3 + 1) != '@') bits |= SBR_ST_INVALID_START;
                                const uint32_t rl = (L - fo) >> 2;
                                if (bits || rl >= A.cap_r) { S.fail = 1; break; }
                                A.seg_rec_off[(size_t)c * A.cap_r + rl] = (L == 3) ? region_lo : S.first[L - 4] + 1;
                                A.seg_key[(size_t)c * A.cap_r + rl] = key;
                                if (key & KEY_NONCANON) atomicOr(&A.hdr->status, SBR_ST_NONCANONICAL);
                                S.wmax[0][0] = rl + 1;  // ascending: any record of warp 0's slices comes later
                                S.wmax[0][1] = p3 + 1;
                                if (p3 + 1 >= region_hi) S.done = 1;
                            }
                            S.synced = 1;
                        } else {
                            S.fail = 1;
                        }
                    }
                }
                __syncthreads();
            }
            if (SBR_ABLATE < 1 && S.synced && fits && !S.fail) {  // (S.fail: some slice of the CTA may have published nothing usable)
                const uint32_t h = S.h, fo = S.first_owned;
                const uint32_t L0 = base + pre_n;  // region-local index of the slice's first newline
                // entries i of this slice that end a record: (h + L0 + i) % 4 == 3.  TWO lanes per record (a 4.5 KB
                // slice ends about fourteen 150-bp records: one lane each would leave the warp half idle): the even
                // lane probes around the header's and the record's final newline and packs the first half of the
                // match window, the odd lane probes around the sequence's newline and packs the second half; one
                // shuffle pair merges them and the even lane commits.  The loop bounds are warp-uniform so that the
                // shuffles run converged.
                constexpr int WPL = KW > 2 ? 2 : 1;  // text words packed per lane
                const uint32_t role = (uint32_t)lane & 1u;
                const uint32_t sbase = smem_u32(sm) - lo;  // shared-window address of batch position 0 (wraps)
                for (uint32_t ib = (3u - (h + L0)) & 3u; ib < n_w; ib += 64u) {
                    const uint32_t i = ib + 4u * ((uint32_t)lane >> 1);
                    uint32_t p3 = 0, p2 = 0, p1 = 0, p0 = 0, start = 0;
                    bool act = i < n_w;
                    if (act) {
                        const uint32_t* e = wl + GHOSTS + i;
                        p3 = e[0]; p2 = e[-1]; p1 = e[-2]; p0 = e[-3];
                        start = e[-4] + 1u;  // (the ghost before region 0 is NOPOS: start 0)
                    }
                    const uint32_t L = L0 + i;
                    // L < fo: started before this region, the previous region finishes it; start >= region_hi: the
                    // next region's record (seen while running over)
                    act = act & (L >= fo) & (start < region_hi);
                    // canonical record (LF ends, bare '+', equal lengths, long enough, next record's '@'): byte probes
                    // from the staged tile, or from the batch text (L2) when the record began in the previous tile;
                    // anything else goes through the general evaluation
                    const uint32_t seq_len = p1 - p0 - 1;
                    const bool shape = act & (p2 == p1 + 2) & (seq_len == p3 - p2 - 1) & (seq_len >= A.bc_len) & (p0 > start);
                    uint32_t codes = 0, iv_ok = 0;
                    if (shape && (start + TAIL_PAD >= lo) & (p3 + 1 < lo + valid)) {  // all of it is staged (tail pad included)
                        const uint32_t qa = role ? p1 : p0, qb = role ? p1 : p3;
                        const uint32_t x = lds_u8(sbase + qa - 1), y = lds_u8(sbase + qb - 1), z = lds_u8(sbase + qb + 1);
                        const uint32_t zc = role ? (uint32_t)'+' : (uint32_t)'@';
                        const uint32_t ok = (x != '\r') & (y != '\r') & (z == zc);
                        if (KW) {
                            const uint64_t pk = pack_prefix_vec<WPL>(sbase + p0 + 1 + 4 * WPL * role, 0xFFFFFFFFu, 0xFFFFu);
                            codes = (uint32_t)pk;
                            iv_ok = (uint32_t)(pk >> 32);
                        }
                        iv_ok |= ok << 16;
                    }
                    const uint32_t codes_o = __shfl_xor_sync(0xFFFFFFFFu, codes, 1);
                    const uint32_t ivok_o = __shfl_xor_sync(0xFFFFFFFFu, iv_ok, 1);
                    if (act && role == 0) {
                        uint64_t key;
                        uint32_t bits = 0;
                        if ((iv_ok & ivok_o) >> 16) {  // shape and probes hold on both lanes
                            const uint32_t cc = (codes | (codes_o << (8 * WPL))) & A.code_mask;
                            const uint32_t iv = ((iv_ok & 0xFFFFu) | ((ivok_o & 0xFFFFu) << (4 * WPL))) & A.inval_mask;
                            key = KW ? ((uint64_t)cc | ((uint64_t)iv << 32)) : (uint64_t)(p0 + 1);
                        } else {
                            key = fastq_slow(sm, A.text, lo, lo + valid, A.nbytes, p0, p1, p2, p3, A.bc_len, A.packed,
                                             (L == 3u && c == 0) ? 1u : 0u);
                            bits = (uint32_t)(key >> 56);
                            key &= (1ull << 56) - 1ull;
                            if (key & KEY_NONCANON) atomicOr(&A.hdr->status, SBR_ST_NONCANONICAL);
                        }
                        const uint32_t rl = (L - fo) >> 2;
                        if (bits || rl >= A.cap_r) S.fail = 1;
                        else {
                            st_keep_u32(seg_off_c + rl, start, keep_policy);
                            st_keep_u64(seg_key_c + rl, key, keep_policy);
                            // records are committed in ascending order: the last one of the slice (or the one that
                            // closes the region) carries the slice's maxima to the CTA
                            const bool closes = p3 + 1 >= region_hi;
                            if (i + 4u >= n_w || closes) {
                                S.wmax[warp][0] = rl + 1;
                                S.wmax[warp][1] = p3 + 1;
                                if (closes) S.done = 1;
                            }
                        }
                    }
                }
            }
        } else {
            // ---- FASTA: every list entry flagged as a record start is one record; its lane evaluates it -------
            const ByteSrc B{sm, A.text, lo, lo + valid};
            const uint32_t rbase = S.rcnt[par] + pre_r + ((c == 0) ? 1u : 0u);  // records before this slice (record 0 of the batch counted)
            const uint32_t sbase = smem_u32(sm) - lo;  // shared-window address of batch position 0 (wraps)
            // lane u takes the u-th flagged entry (every other newline of single-line FASTA is a start: going chunk by
            // chunk would leave half of the lanes idle in twice as many passes)
            for (uint32_t u = (uint32_t)lane; fits && u < nr_w; u += 32u) {
                const uint32_t idx = S.ridx[warp][u];
                {
                    const uint32_t p = wl[GHOSTS + idx] & POS_MASK;
                    // the two newlines behind the record start (the header's, the first sequence line's): from the own
                    // list, or from the head the next slice published
                    auto ahead = [&](uint32_t j) -> uint32_t {
                        if (j < n_w) return wl[GHOSTS + j] & POS_MASK;
                        return (warp + 1 < FAST_WARPS && j - n_w < 2) ? S.whead[par][warp + 1][j - n_w] : NOPOS;
                    };
                    const uint32_t hend = ahead(idx + 1);  // the header's newline
                    uint64_t key = 0;
                    uint32_t why = 3;  // 3: not decided by the fast path
                    // single-line prefix made of A/C/G/T only, staged: pack it straight from shared memory (a newline,
                    // CR, N ... inside the window shows up as an invalid base and sends the record to the general path)
                    if (KW && hend != NOPOS && hend + 1 + 4 * (KW ? KW : 1) + 8 < lo + valid) {
                        key = pack_prefix_vec<(KW ? KW : 1)>(sbase + hend + 1, A.code_mask, A.inval_mask);
                        // all A/C/G/T: no line end can hide in the window.  Otherwise (an N, say) the window must be
                        // known to lie on the first sequence line: the next newline is listed and far enough
                        if ((key >> 32) == 0) why = 0;
                        else if (lds_u8(sbase + hend + 1) != '>') {  // ('>' right behind the header: an empty record, the general path's case)
                            // (+1: the line may end in CR LF, and the CR must stay outside the window too)
                            const uint32_t e2 = ahead(idx + 2);
                            if (e2 != NOPOS && e2 > hend + A.bc_len + 1) why = 0;
                        }
                    }
                    if (why == 3) why = fasta_eval(B, A.nbytes, p + 1, hend, A.bc_len, A.packed, key);
                    const uint32_t rl = rbase + u;
                    // a prefix cut by the end of a non-final batch belongs to the incomplete last record: no error
                    if (rl >= A.cap_r || why == 1 || (why == 2 && A.final_batch)) S.fail = 1;
                    else {
                        st_keep_u32(seg_off_c + rl, p + 1, keep_policy);
                        st_keep_u64(seg_key_c + rl, key, keep_policy);
                    }
                    if (u == nr_w - 1)  // ascending over the tiles: a plain store per warp, the maximum is taken at the end
                        S.wlast[warp] = ((unsigned long long)(base + pre_n + idx + 1) << 32) | (unsigned long long)(p + 1);
                }
            }
            if (tid == 0 && c == 0 && it == 0) {
                uint64_t key;
                if (B(0) != '>') S.fail = 1;
                const uint32_t why = fasta_eval(B, A.nbytes, 0u, (fits && n_w > 0) ? (wl[GHOSTS] & POS_MASK) : NOPOS, A.bc_len, A.packed, key);
                if (why == 1 || (why == 2 && A.final_batch)) S.fail = 1;
                A.seg_rec_off[0] = 0;
                A.seg_key[0] = key;
            }
        }
        if (FAST_STAGES == 1) {
            __syncthreads();  // everybody is done with the tile's text (and S.done / S.synced are final)
            if (warp == 0) {
                if (FMT == SBR_FMT_FASTQ) {  // the tile's last TAIL_PAD bytes -> the pad in front of the stage
                    const uint4 v = *reinterpret_cast<const uint4*>(sm + FAST_TILE - TAIL_PAD + 16 * lane);
                    __syncwarp();
                    *reinterpret_cast<uint4*>(stage_buf - TAIL_PAD + 16 * lane) = v;
                    __syncwarp();
                }
                if (lane == 0 && tile + 1 < t_hi) issue(0, tile + 1);
            }
        }
    }
    __syncthreads();

    if (tid == 0) {
        const uint32_t n_it = idle ? 0u : t_hi - t_lo;  // own tiles processed: the final counts sit in parity n_it & 1
        RegionInfo e;
        e.nl_own = S.nl_own;
        uint32_t nrec = 0, last_end = 0;
        for (int w = 0; w < FAST_WARPS; ++w) { nrec = max(nrec, S.wmax[w][0]); last_end = max(last_end, S.wmax[w][1]); }
        e.nrec = (FMT == SBR_FMT_FASTA) ? S.rcnt[n_it & 1u] + (c == 0 ? 1u : 0u) : nrec;
        e.phase = (FMT == SBR_FMT_FASTQ && S.synced) ? S.h : PHASE_UNKNOWN;
        e.flags = S.fail ? 1u : 0u;
        unsigned long long last_start = 0;
        for (int w = 0; w < FAST_WARPS; ++w) last_start = max(last_start, S.wlast[w]);
        e.last_start_nl = (uint32_t)(last_start >> 32);
        e.last_start_off = (uint32_t)last_start;
        e.pad[0] = e.pad[1] = 0;
        A.info[c] = e;
        if (last_end) atomicMax(&A.hdr->consumed, last_end);
        if (S.crlf) atomicOr(&A.hdr->status, SBR_ST_NONCANONICAL);
        // the ticket: everything this CTA wrote (the barrier above ordered its other threads' stores and atomics
        // before this fence) is visible to whoever draws the last number
        __threadfence();
        S.is_last = atomicAdd(&A.hdr->arrive, 1u) == gridDim.x - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (!S.is_last) return;
    __threadfence();
    scan_finish<FMT>(A, gridDim.x, reinterpret_cast<uint32_t*>(dyn_smem));  // the stage is free: scratch
}

}  // namespace

uint32_t scan_fast_tile_bytes() { return (uint32_t)FAST_TILE; }

size_t scan_fast_smem_bytes(uint32_t fmt) {
    return (size_t)FAST_STAGES * stage_stride((int)fmt) +
           (fmt == SBR_FMT_FASTA ? sizeof(FastSmem<SBR_FMT_FASTA>) : sizeof(FastSmem<SBR_FMT_FASTQ>));
}

using FastKernel = void (*)(FastArgs);
// the kernels are specialised on the number of text words covering the match window
template <int FMT>
static FastKernel fast_kernel(uint32_t bc_len, uint32_t packed) {
    if (!packed) return k_scan_fast<FMT, 0>;
    switch ((bc_len + 3) / 4) {
        case 1: return k_scan_fast<FMT, 1>;
        case 2: return k_scan_fast<FMT, 2>;
        case 3: return k_scan_fast<FMT, 3>;
        default: return k_scan_fast<FMT, 4>;
    }
}

cudaError_t scan_fast_prepare(int device, int* grid_fastq, int* grid_fasta) {
    cudaError_t e;
    int sms = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return e;
    int resident[2] = {1 << 30, 1 << 30};  // FASTA, FASTQ: one grid for every variant of a format, the smallest residency
    for (uint32_t fmt = SBR_FMT_FASTA; fmt <= SBR_FMT_FASTQ; ++fmt)
        for (uint32_t kw = 0; kw <= 4; ++kw) {
            const FastKernel k = fmt == SBR_FMT_FASTA ? fast_kernel<SBR_FMT_FASTA>(4 * kw, kw ? 1u : 0u)
                                                     : fast_kernel<SBR_FMT_FASTQ>(4 * kw, kw ? 1u : 0u);
            if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_fast_smem_bytes(fmt))) !=
                cudaSuccess)
                return e;
            // the SM's whole 228 KB as shared memory: four CTAs of the scan need ~193 KB, and the match kernel of the
            // previous batch must find room beside them (with the 196 KB split the driver would pick it cannot)
            int carve = cudaSharedmemCarveoutMaxShared;
            if (const char* ev = getenv("SBR_SCAN_CARVEOUT")) carve = atoi(ev);
            if ((e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve)) != cudaSuccess) return e;
            int r = 0;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, k, FAST_THREADS, scan_fast_smem_bytes(fmt))) != cudaSuccess)
                return e;
            int& best = resident[fmt - SBR_FMT_FASTA];
            best = r < best ? r : best;
        }
    const int a = (resident[0] < 1 ? 1 : resident[0]) * sms, q = (resident[1] < 1 ? 1 : resident[1]) * sms;
    *grid_fastq = q > MAX_REGIONS ? MAX_REGIONS : q;
    *grid_fasta = a > MAX_REGIONS ? MAX_REGIONS : a;
    return cudaSuccess;
}

// Regions of a batch: at most max_regions CTAs, at least 4 tiles (96 KB) each so that the work at a
// region's start and end is amortised and every region sees enough newlines to find its phase; the
// tiles that do not divide evenly go one each to the first regions.
uint32_t scan_fast_regions(uint32_t nbytes, int max_regions, uint32_t* tiles_per_region, uint32_t* extra_regions) {
    const uint32_t ntiles = (nbytes + FAST_TILE - 1) / FAST_TILE;
    const uint32_t min_tiles = 4;
    uint32_t R = ntiles / min_tiles;
    if (R > (uint32_t)max_regions) R = (uint32_t)max_regions;
    if (R == 0) R = 1;
    *tiles_per_region = ntiles / R;
    *extra_regions = ntiles % R;
    return R;
}

cudaError_t scan_fast_launch(uint32_t fmt, const uint8_t* d_text, uint32_t nbytes, uint32_t max_records, uint32_t cap_r,
                             uint32_t n_regions, uint32_t tiles_per_region, uint32_t extra_regions, uint32_t bc_len,
                             uint32_t packed,
                             uint32_t final_batch, BatchHeader* hdr, RegionInfo* info, uint32_t* region_base,
                             uint32_t* seg_rec_off, uint64_t* seg_key, uint32_t* rec_off_dense, uint4* lookback_desc,
                             uint32_t n_desc, uint32_t* stats, cudaStream_t stream) {
    FastArgs A;
    A.text = d_text;
    A.nbytes = nbytes;
    A.ntiles = (nbytes + FAST_TILE - 1) / FAST_TILE;
    A.tiles_per_region = tiles_per_region;
    A.extra_regions = extra_regions;
    A.cap_r = cap_r;
    A.bc_len = bc_len;
    A.packed = packed;
    A.final_batch = final_batch;
    A.c7f = 0x7F7F7F7Fu;
    A.c0a = 0x0A0A0A0Au;
    A.c80 = 0x80808080u;
    A.c3e = 0x3E3E3E3Eu;
    A.code_mask = bc_len >= 16 ? 0xFFFFFFFFu : ((1u << (2 * bc_len)) - 1u);
    A.inval_mask = bc_len >= 16 ? 0xFFFFu : ((1u << bc_len) - 1u);
    A.hdr = hdr;
    A.info = info;
    A.seg_rec_off = seg_rec_off;
    A.seg_key = seg_key;
    A.max_records = max_records;
    A.n_desc = n_desc;
    A.region_base = region_base;
    A.rec_off_dense = rec_off_dense;
    A.lookback_desc = lookback_desc;
    A.stats = stats;
    size_t smem = scan_fast_smem_bytes(fmt);
    if (fmt == SBR_FMT_FASTQ) fast_kernel<SBR_FMT_FASTQ>(bc_len, packed)<<<n_regions, FAST_THREADS, smem, stream>>>(A);
    else fast_kernel<SBR_FMT_FASTA>(bc_len, packed)<<<n_regions, FAST_THREADS, smem, stream>>>(A);
    return cudaGetLastError();
}

}  // namespace sbr
