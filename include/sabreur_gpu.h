/*
 * sabreur_gpu.h -- C ABI of the B200-native demultiplexing hot path (libsabreur_gpu.so).
 *
 * The reference (Ebedthan/sabreur v0.7.0, pure Rust) has no FFI seam: its hot path is the Rust
 * function boundary main -> se_demux / pe_demux.  This header is the seam a host would bind instead:
 *
 *   replaces                                        reference file:line
 *   ----------------------------------------------  -------------------------------------------
 *   record source (needletail parse_fastx_reader)   src/demux.rs:23,49 / 85-86,101,114
 *   bc_cmp + iter().find() first-match loop         src/utils.rs:115-123, src/demux.rs:52-54,103,116
 *   table prep (bc_len, candidate order)            src/demux.rs:31-37, 89-92
 *   routing of each record to its barcode's stream  src/demux.rs:58-59,62-63,104-108,117-121
 *   per-barcode counters nb_records                 src/demux.rs:58,104,117
 *
 * Everything is plain C: pointers and sizes only, no C++/torch types.  All functions return 0
 * (SBR_OK) or a negative sbr_status; nothing throws or aborts across the boundary.
 * A context is bound to one GPU and is NOT thread-safe: one host thread per context; distinct
 * contexts (GPUs) are independent.  The library owns every buffer it hands out.
 */
#ifndef SABREUR_GPU_H
#define SABREUR_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBR_ABI_VERSION 2

typedef enum {
    SBR_OK = 0,
    SBR_E_INVALID_ARG = -1,
    SBR_E_CUDA = -2,        /* a CUDA runtime call failed: see sbr_last_error() */
    SBR_E_NO_DEVICE = -3,   /* no usable CUDA device: there is NO CPU fallback */
    SBR_E_UNSUPPORTED = -4,
    SBR_E_BUSY = -5,        /* slot submitted and not yet waited for */
    SBR_E_NOMEM = -6
} sbr_status;

/* sbr_result.status_flags: what the record scan found in a batch (parse violations are reported,
 * not fatal, so the host can reproduce the reference's behaviour: SE aborts on the first bad record
 * (src/demux.rs:50), PE stops silently (src/demux.rs:101,114), a short read panics (demux.rs:54)). */
#define SBR_ST_INVALID_START 0x01u  /* record does not start with '@' / '>' */
#define SBR_ST_INVALID_SEP   0x02u  /* FASTQ third line does not start with '+' */
#define SBR_ST_UNEQUAL_LEN   0x04u  /* FASTQ len(seq) != len(qual) */
#define SBR_ST_SHORT_READ    0x08u  /* len(seq) < bc_len */
#define SBR_ST_TRUNCATED     0x10u  /* final batch ends inside a record */
#define SBR_ST_REC_OVERFLOW  0x20u  /* more than max_records records: `consumed` covers those kept */
#define SBR_ST_NONCANONICAL  0x40u  /* informational: some record's canonical output bytes differ
                                       from its input bytes (CR, "+id" line, multi-line FASTA) */
#define SBR_ST_ERROR_MASK    0x1Fu

#define SBR_FMT_AUTO  0u
#define SBR_FMT_FASTA 1u
#define SBR_FMT_FASTQ 2u

/* sbr_config.flags */
#define SBR_CFG_NO_LUT   0x1u  /* force the XOR+popc table walk even where a first-match LUT fits */
#define SBR_CFG_GENERIC  0x2u  /* force the byte-exact match kernel (it is selected automatically when
                                  the table holds non-ACGT bytes or bc_len > 16) */
#define SBR_CFG_TIMING   0x4u  /* record CUDA events around every kernel: see sbr_timing_collect() */
#define SBR_CFG_EXACT_SCAN 0x8u /* skip the region-parallel scan: always run the look-back scan */
#define SBR_CFG_NO_FUSE 0x10u   /* run match and partition as separate launches (default: one cooperative launch) */
#define SBR_CFG_OVERLAP 0x20u   /* device-resident batches (sbr_demux_device): match + partition of a batch run on a
                                   stream of the slot's own, beside the NEXT batch's record scan on the caller's stream;
                                   see sbr_device_join() */

typedef struct sbr_ctx sbr_ctx;

typedef struct {
    int32_t device;          /* CUDA device ordinal */
    uint32_t fmt;            /* SBR_FMT_*; AUTO = decided from the first byte of each batch */
    uint32_t n_barcodes;     /* rows of the barcode file, FILE ORDER (= candidate order) */
    uint32_t bc_len;         /* row stride in `barcodes`, and the match window: seq[..bc_len] */
    const uint8_t* barcodes; /* n_barcodes * bc_len raw ASCII bytes */
    const uint32_t* bc_lens; /* optional: significant bytes per row (<= bc_len); NULL = all bc_len */
    uint32_t mismatch;       /* --mismatch */
    uint32_t n_slots;        /* pipeline depth (>= 1; 2+ to overlap copies with kernels) */
    uint64_t batch_bytes;    /* capacity of one slot's text buffer, <= 1 GiB */
    uint32_t max_records;    /* record capacity per batch; 0 = batch_bytes / 32 */
    uint32_t flags;          /* SBR_CFG_* */
} sbr_config;

typedef struct {
    uint64_t batch_seq;        /* echoed from sbr_submit */
    uint64_t consumed;         /* bytes that formed whole records; the tail must be re-fed */
    uint32_t n_records;
    uint32_t status_flags;     /* SBR_ST_* */
    uint32_t first_bad_record; /* valid when status_flags & SBR_ST_ERROR_MASK */
    uint32_t first_bad_flags;  /* the SBR_ST_* bits of that record */
    uint32_t fmt;              /* SBR_FMT_FASTA / SBR_FMT_FASTQ as scanned */
    uint32_t n_bins;           /* n_barcodes + 1; bin n_barcodes = unknown */
    const uint32_t* rec_off;   /* [n_records + 1] batch-relative record starts */
    const uint16_t* assign;    /* [n_records] barcode row, n_barcodes == unknown */
    const uint32_t* bin_count; /* [n_bins] */
    const uint32_t* bin_start; /* [n_bins + 1] exclusive scan of bin_count */
    const uint32_t* order;     /* [n_records] record indices grouped by bin, ascending within a bin */
    const uint64_t* rec_key;   /* [n_records] per-record scan word (packed prefix + flags), see DESIGN.md;
                                  NULL unless status_flags has SBR_ST_NONCANONICAL */
    uint32_t scan_path;        /* 1: the region-parallel scan's result stood, 0: the look-back scan ran */
    uint32_t reserved;
} sbr_result;

/* ---- life cycle ------------------------------------------------------------------------------ */
int sbr_abi_version(void);
int sbr_device_count(void); /* >= 0, or a negative sbr_status */
int sbr_create(sbr_ctx** out, const sbr_config* cfg);
void sbr_destroy(sbr_ctx* ctx);
const char* sbr_last_error(const sbr_ctx* ctx); /* ctx may be NULL: last create() failure */

/* ---- streaming path: pinned host slots -> GPU -> pinned results --------------------------------
 * sbr_slot_buffer: the pinned host buffer of a slot.  *capacity = batch_bytes is what the caller may fill; the
 * allocation is 64 bytes larger, and that padding belongs to the library (final-newline patch, 16-byte
 * zero fill).  The host decompressor writes text straight into the buffer.
 * sbr_submit: asynchronous H2D copy -> record scan -> match -> partition -> D2H of the results.
 * The batch is host_ptr[offset .. offset+nbytes) (offset + nbytes <= capacity): a host that reserves
 * room at the front of a slot can prepend the unconsumed tail of the previous batch without moving
 * the freshly read data.  rec_off / consumed are relative to `offset`.
 * Batches must start on a record boundary.  final_batch != 0: the batch is the end of the stream
 * (a missing final newline is supplied, up to three trailing blank lines are dropped).
 * sbr_wait: blocks until the slot's batch is done; pointers in *res stay valid until the slot is
 * submitted again. */
int sbr_slot_buffer(sbr_ctx* ctx, uint32_t slot, uint8_t** host_ptr, uint64_t* capacity);
int sbr_submit(sbr_ctx* ctx, uint32_t slot, uint64_t offset, uint64_t nbytes, uint64_t batch_seq,
               int final_batch);
int sbr_wait(sbr_ctx* ctx, uint32_t slot, sbr_result* res);

/* ---- device-resident path (kernel tests, HBM-resident benchmark) -------------------------------
 * Runs scan -> match -> partition on text that already lives in device memory (16-byte aligned),
 * on `stream` (a cudaStream_t passed as void*; NULL = the slot's own stream), using the slot's
 * device workspace.  Nothing is copied by the call itself: sbr_slot_device_result() exposes the
 * DEVICE pointers, and a following sbr_wait() on the slot copies the results to pinned host memory.
 * d_text must be readable up to the next multiple of 16 bytes past nbytes.  The device results are final
 * only after sbr_wait(): a batch the region-parallel scan hands over (malformed, newline-dense ...) reads
 * "no records" until sbr_wait() has re-run it through the exact scan. */
int sbr_demux_device(sbr_ctx* ctx, uint32_t slot, const uint8_t* d_text, uint64_t nbytes,
                     int final_batch, void* stream);
int sbr_slot_device_result(sbr_ctx* ctx, uint32_t slot, sbr_result* res_device_ptrs);
/* SBR_CFG_OVERLAP: makes `stream` wait for the match + partition work still outstanding on the slots' own
 * streams (one cudaStreamWaitEvent per slot; nothing blocks on the host).  Call it before recording the event
 * that ends a timed region, or before work on `stream` that reads the device results.  A no-op otherwise. */
int sbr_device_join(sbr_ctx* ctx, void* stream);
/* Device-side counters of the region-parallel scan since the context was created (or last reset):
 * out[0] = batches it finished, out[1] = batches it handed over to the exact look-back scan (these read
 * "no records" on the device until sbr_wait() has re-run them).  Synchronises the device. */
int sbr_scan_stats(sbr_ctx* ctx, uint64_t out[2], int reset);
/* Geometry of the region-parallel scan for `fmt` (SBR_FMT_FASTA / SBR_FMT_FASTQ): a batch is cut into at most *regions
 * contiguous regions (one per resident CTA) of whole tiles of *tile_bytes.  A caller that is free to choose its batch size
 * gets evenly loaded regions from batches of (a multiple of) regions * tile_bytes bytes, or a little less. */
int sbr_scan_geometry(const sbr_ctx* ctx, uint32_t fmt, uint32_t* tile_bytes, uint32_t* regions);

/* The first-match lookup table built from the barcode table (debug / test access).
 * Returns the number of entries (4^bc_len) or 0 when the context does not use a LUT. */
uint64_t sbr_lut_entries(const sbr_ctx* ctx);
int sbr_lut_copy(sbr_ctx* ctx, uint16_t* host_out, uint64_t n_entries);

/* Debug / test access to the region table of the slot's last region-parallel scan: eight uint32 per
 * region {newlines, records, detected phase, flags, last_start_nl, last_start_off, 0, 0}.
 * Returns the number of regions copied (<= max_regions) or a negative sbr_status. */
int sbr_debug_regions(sbr_ctx* ctx, uint32_t slot, uint32_t* host_out, uint32_t max_regions);

/* Per-kernel device time accumulated since the last call (SBR_CFG_TIMING): ms[0]=scan,
 * ms[1]=match, ms[2]=partition; *launches = kernel launches covered.  Synchronises the device. */
int sbr_timing_collect(sbr_ctx* ctx, double ms[3], uint64_t* batches, uint64_t* launches);

/* ---- synthetic workload (SURVEY 8d / DESIGN.md "Synthetic generator") --------------------------
 * Counter-based and position-independent: record i is a pure function of (spec, i); the device
 * and host generators produce identical bytes. */
typedef struct {
    uint64_t seed;        /* read stream seed */
    uint32_t fmt;         /* SBR_FMT_FASTA / SBR_FMT_FASTQ */
    uint32_t read_len;    /* bases per read */
    uint32_t mate;        /* 0 -> "/1", 1 -> "/2" (FASTQ header suffix; also mixed into the stream) */
    uint32_t n_barcodes;
    uint32_t bc_len;
    uint32_t reserved;
    const uint8_t* barcodes; /* n_barcodes * bc_len ASCII (host pointer) */
} sbr_synth_spec;

/* greedy min-distance barcode table: row j accepted iff Hamming distance to all accepted rows
 * >= min_dist.  Returns SBR_OK or SBR_E_UNSUPPORTED if the table cannot be filled. */
int sbr_synth_table(uint64_t seed, uint32_t n_barcodes, uint32_t bc_len, uint32_t min_dist,
                    uint8_t* out /* n_barcodes * bc_len */);
uint32_t sbr_synth_record_bytes(const sbr_synth_spec* spec);
int sbr_synth_host(const sbr_synth_spec* spec, uint64_t first_record, uint64_t n_records,
                   uint8_t* out);
int sbr_synth_device(int device, const sbr_synth_spec* spec, uint64_t first_record,
                     uint64_t n_records, uint8_t* d_out, void* stream);

/* ---- plain device memory helpers (so hosts without a CUDA binding can stage data) -------------- */
int sbr_dev_alloc(int device, uint64_t nbytes, void** d_ptr);
int sbr_dev_free(int device, void* d_ptr);
int sbr_dev_upload(int device, void* d_dst, const void* h_src, uint64_t nbytes);
int sbr_dev_download(int device, void* h_dst, const void* d_src, uint64_t nbytes);
int sbr_host_alloc_pinned(uint64_t nbytes, void** h_ptr);
int sbr_host_free_pinned(void* h_ptr);

#ifdef __cplusplus
}
#endif
#endif /* SABREUR_GPU_H */
