/*
 * sabreur_oracle.h -- CPU restatement of sabreur's demultiplexing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the
 * product (sabreur_b200/, include/).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it, and only as the checker or
 * as the timed CPU baseline -- never as a fallback for the CUDA path.
 *
 * Parity status: PARTIALLY PINNED.  The reference (Rust, /root/reference) cannot be built in
 * this environment (no cargo/rustc) and its parser/serialiser live in un-vendored crates
 * (needletail 0.5.1, niffler 2.5.0 -- Cargo.lock:518-544).  What IS pinned against the
 * reference's own tests: bc_cmp truth table (src/utils.rs:177-205) and split_by_tab
 * (src/utils.rs:207-221).  Everything else follows the reference source line by line
 * (citations below) under the deterministic refinement documented in DESIGN.md
 * (barcode-file row order as candidate order, bc_len from row 0, "XXX" sentinel last).
 */
#ifndef SABREUR_ORACLE_H
#define SABREUR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_FMT_FASTA = 1, ORC_FMT_FASTQ = 2 };

enum {
    ORC_OK = 0,
    ORC_E_EMPTY = 1,          /* empty input (needletail: EmptyFile) */
    ORC_E_INVALID_START = 2,  /* record does not start with '>' / '@' */
    ORC_E_INVALID_SEP = 3,    /* FASTQ third line does not start with '+' */
    ORC_E_UNEQUAL_LEN = 4,    /* FASTQ len(seq) != len(qual) */
    ORC_E_TRUNCATED = 5,      /* final batch ends inside a record */
    ORC_E_SHORT_READ = 6,     /* len(seq) < bc_len: reference panics, src/demux.rs:54 */
    ORC_E_CAPACITY = 7
};

enum { ORC_REC_NONCANONICAL = 1u }; /* canonical output bytes differ from the input bytes */

typedef struct {
    uint64_t rec_off;  /* first byte of the record ('@' or '>') */
    uint64_t rec_end;  /* one past the last byte of the record (incl. its final '\n' if any) */
    uint64_t id_off, id_len;   /* header line without marker, trailing CR removed */
    uint64_t seq_off, seq_end; /* raw span of the sequence line(s), trailing CR/LF removed */
    uint64_t qual_off, qual_len; /* FASTQ only */
    uint32_t flags;
    uint32_t pad;
} orc_rec;

/* src/utils.rs:115-123 -- zip => min(len) bytes, u8 accumulator (wraps), <= mismatch */
int orc_bc_cmp(const uint8_t* bc, size_t bc_len, const uint8_t* seq, size_t seq_len,
               uint8_t mismatch);

/* needletail::parse_fastx_reader + next() as used at src/demux.rs:23,49,85-86,101,114.
 * Parses whole records from text[0..n).  final_batch != 0: the last record may lack its
 * final '\n' and up to three trailing blank lines are tolerated; otherwise parsing stops
 * at the last complete record and *consumed says how many bytes formed whole records.
 * Returns the number of records written, or -1 on error (*err, *err_rec say which). */
long orc_parse(const uint8_t* text, size_t n, int final_batch, orc_rec* recs, size_t cap,
               int* fmt, size_t* consumed, int* err, size_t* err_rec);

/* record.seq(): FASTA lines joined, CR/LF removed.  Writes at most cap bytes, returns the
 * full cleaned length. */
size_t orc_seq(const uint8_t* text, const orc_rec* r, int fmt, uint8_t* out, size_t cap);

/* src/demux.rs:52-54 / 101-103 / 114-116 with candidate order = table row order:
 * assign[i] = first row b with bc_cmp(barcode[b], seq[..bc_len], m), else n_barcodes.
 * bc_len = length of row 0 (src/demux.rs:31-34, 90 under refinement H2).
 * Barcodes are rows of `stride` bytes, row b has bc_lens[b] significant bytes.
 * Returns ORC_OK or ORC_E_SHORT_READ (with *err_rec = first short record). */
int orc_assign(const uint8_t* text, const orc_rec* recs, size_t n_rec, int fmt,
               const uint8_t* barcodes, const uint32_t* bc_lens, uint32_t n_barcodes,
               uint32_t stride, uint8_t mismatch, uint16_t* assign, size_t* err_rec);

/* the partition implied by src/demux.rs:58-59,62-63: per-bin counts, exclusive starts and
 * the stable (input-order) list of record indices per bin.  nbins = n_barcodes + 1. */
void orc_partition(const uint16_t* assign, size_t n_rec, uint32_t nbins, uint32_t* bin_count,
                   uint32_t* bin_start /* nbins+1 */, uint32_t* order);

/* src/utils.rs:133-158 + needletail write_fasta/write_fastq with LineEnding::Unix:
 * ">id\nseq\n" / "@id\nseq\n+\nqual\n".  Returns bytes written (out may be NULL to size). */
size_t orc_serialize(const uint8_t* text, const orc_rec* r, int fmt, uint8_t* out);

/* Reference-faithful streaming loop (src/demux.rs:49-66): one record at a time, byte-wise
 * bc_cmp over the table in order (+ the "XXX" sentinel compare on a miss), canonical bytes
 * appended to the bin's in-memory sink (stands in for the per-record write_seqs, without the
 * syscalls).  Used as the timed CPU baseline.  Returns records processed or -1 on error.
 * sink_bytes[nbins] receives the bytes appended per bin, bin_count[nbins] the hits. */
long orc_demux_stream(const uint8_t* text, size_t n, const uint8_t* barcodes,
                      const uint32_t* bc_lens, uint32_t n_barcodes, uint32_t stride,
                      uint8_t mismatch, uint64_t* bin_count, uint64_t* sink_bytes,
                      uint64_t* sink_hash, int keep_bytes);

/* FNV-1a 64 of a buffer: used to compare per-bin output streams cheaply. */
uint64_t orc_fnv1a(const uint8_t* p, size_t n, uint64_t h);

#ifdef __cplusplus
}
#endif
#endif
