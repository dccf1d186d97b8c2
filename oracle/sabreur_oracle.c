/*
 * sabreur_oracle.c -- CPU restatement of sabreur's demux hot path (see sabreur_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into, imported by, or executed from the product.
 *
 * Follows /root/reference (Ebedthan/sabreur v0.7.0):
 *   bc_cmp ................ src/utils.rs:115-123
 *   SE loop ............... src/demux.rs:49-66
 *   PE loops .............. src/demux.rs:101-110, 114-123 (mates are independent streams)
 *   table prep ............ src/demux.rs:31-37, 89-92
 *   serialisation ......... src/utils.rs:133-158 (needletail write_fasta / write_fastq, Unix)
 * Record parsing restates the observable behaviour of needletail 0.5.1 (un-vendored crate,
 * Cargo.lock:518-529): format by first byte; FASTQ strictly four lines per record; FASTA
 * record = '>' at start of line up to the next "\n>"; CR before LF stripped.
 */
#include "sabreur_oracle.h"

#include <stdlib.h>
#include <string.h>

int orc_bc_cmp(const uint8_t* bc, size_t bc_len, const uint8_t* seq, size_t seq_len,
               uint8_t mismatch) {
    /* bc.iter().zip(seq.iter()).map(|(a,b)| (a != b) as u8).sum::<u8>() <= mismatch
     * release profile has overflow-checks off (Cargo.toml:29-35): the u8 sum wraps. */
    size_t n = bc_len < seq_len ? bc_len : seq_len;
    uint8_t acc = 0;
    for (size_t i = 0; i < n; ++i) acc = (uint8_t)(acc + (uint8_t)(bc[i] != seq[i]));
    return acc <= mismatch;
}

static const uint8_t* find_nl(const uint8_t* p, const uint8_t* end) {
    return p < end ? (const uint8_t*)memchr(p, '\n', (size_t)(end - p)) : NULL;
}

/* true iff every line of [p,end) is empty after CR trimming */
static int blank_tail(const uint8_t* p, const uint8_t* end) {
    for (; p < end; ++p)
        if (*p != '\n' && *p != '\r') return 0;
    return 1;
}

static long parse_fastq(const uint8_t* text, size_t n, int final_batch, orc_rec* recs,
                        size_t cap, size_t* consumed, int* err, size_t* err_rec) {
    const uint8_t* end = text + n;
    const uint8_t* pos = text;
    size_t nr = 0;
    *err = ORC_OK;
    while (pos < end) {
        if (nr >= cap) {
            if (final_batch && blank_tail(pos, end)) { pos = end; break; }
            *err = ORC_E_CAPACITY; *err_rec = nr; *consumed = (size_t)(pos - text); return -1;
        }
        const uint8_t* p0 = find_nl(pos, end);
        const uint8_t* p1 = p0 ? find_nl(p0 + 1, end) : NULL;
        const uint8_t* p2 = p1 ? find_nl(p1 + 1, end) : NULL;
        const uint8_t* p3 = p2 ? find_nl(p2 + 1, end) : NULL;
        int virtual_nl = 0;
        if (!p3) {
            if (!final_batch) break; /* incomplete tail: host re-feeds it */
            if (blank_tail(pos, end)) { pos = end; break; } /* <=3 trailing blank lines */
            if (!p2) { *err = ORC_E_TRUNCATED; *err_rec = nr; *consumed = (size_t)(pos - text); return -1; }
            p3 = end; /* last record without its final newline */
            virtual_nl = 1;
        }
        orc_rec r;
        memset(&r, 0, sizeof r);
        r.rec_off = (uint64_t)(pos - text);
        r.rec_end = (uint64_t)((virtual_nl ? p3 : p3 + 1) - text);
        int cr0 = (p0 > pos && p0[-1] == '\r');
        int cr1 = (p1 > p0 + 1 && p1[-1] == '\r');
        int cr2 = (p2 > p1 + 1 && p2[-1] == '\r');
        int cr3 = (p3 > p2 + 1 && p3[-1] == '\r');
        if (*pos != '@') { *err = ORC_E_INVALID_START; *err_rec = nr; *consumed = r.rec_off; return -1; }
        if (!(p1 + 1 < p2 && p1[1] == '+')) { /* empty third line or not starting with '+' */
            *err = ORC_E_INVALID_SEP; *err_rec = nr; *consumed = r.rec_off; return -1;
        }
        r.id_off = r.rec_off + 1;
        r.id_len = (uint64_t)((p0 - cr0) - (pos + 1));
        r.seq_off = (uint64_t)(p0 + 1 - text);
        r.seq_end = (uint64_t)((p1 - cr1) - text);
        r.qual_off = (uint64_t)(p2 + 1 - text);
        r.qual_len = (uint64_t)((p3 - cr3) - (p2 + 1));
        if (r.seq_end - r.seq_off != r.qual_len) { *err = ORC_E_UNEQUAL_LEN; *err_rec = nr; *consumed = r.rec_off; return -1; }
        int sep_bare = ((p2 - cr2) - (p1 + 1)) == 1;
        if (cr0 || cr1 || cr2 || cr3 || !sep_bare || virtual_nl) r.flags |= ORC_REC_NONCANONICAL;
        recs[nr++] = r;
        pos = virtual_nl ? end : p3 + 1;
    }
    *consumed = (size_t)(pos - text);
    return (long)nr;
}

static long parse_fasta(const uint8_t* text, size_t n, int final_batch, orc_rec* recs,
                        size_t cap, size_t* consumed, int* err, size_t* err_rec) {
    const uint8_t* end = text + n;
    const uint8_t* pos = text;
    size_t nr = 0;
    *err = ORC_OK;
    while (pos < end) {
        if (nr >= cap) { *err = ORC_E_CAPACITY; *err_rec = nr; *consumed = (size_t)(pos - text); return -1; }
        if (*pos != '>') { *err = ORC_E_INVALID_START; *err_rec = nr; *consumed = (size_t)(pos - text); return -1; }
        const uint8_t* h = find_nl(pos, end);
        const uint8_t* next = NULL; /* start of the next record */
        if (h) {
            const uint8_t* q = h;
            while (q) {
                if (q + 1 < end && q[1] == '>') { next = q + 1; break; }
                q = find_nl(q + 1, end);
            }
        }
        if (!next) {
            if (!final_batch) break; /* cannot know where the record ends yet */
            next = end;
        }
        orc_rec r;
        memset(&r, 0, sizeof r);
        r.rec_off = (uint64_t)(pos - text);
        r.rec_end = (uint64_t)(next - text);
        const uint8_t* hend = h ? h : end; /* header-only record at EOF */
        int cr0 = (hend > pos + 1 && hend[-1] == '\r');
        r.id_off = r.rec_off + 1;
        r.id_len = (uint64_t)((hend - cr0) - (pos + 1));
        const uint8_t* s0 = h ? h + 1 : end;
        const uint8_t* s1 = next;
        while (s1 > s0 && (s1[-1] == '\n' || s1[-1] == '\r')) --s1;
        r.seq_off = (uint64_t)(s0 - text);
        r.seq_end = (uint64_t)(s1 - text);
        int canonical = h && !cr0 && (next - s1) == 1; /* exactly one '\n' after the sequence */
        if (canonical)
            for (const uint8_t* q = s0; q < s1; ++q)
                if (*q == '\n' || *q == '\r') { canonical = 0; break; }
        if (!canonical) r.flags |= ORC_REC_NONCANONICAL;
        recs[nr++] = r;
        pos = next;
    }
    *consumed = (size_t)(pos - text);
    return (long)nr;
}

long orc_parse(const uint8_t* text, size_t n, int final_batch, orc_rec* recs, size_t cap,
               int* fmt, size_t* consumed, int* err, size_t* err_rec) {
    *consumed = 0;
    *err_rec = 0;
    if (n == 0) { *err = ORC_E_EMPTY; return -1; }
    if (*fmt == 0) {
        if (text[0] == '>') *fmt = ORC_FMT_FASTA;
        else if (text[0] == '@') *fmt = ORC_FMT_FASTQ;
        else { *err = ORC_E_INVALID_START; return -1; }
    }
    if (*fmt == ORC_FMT_FASTQ)
        return parse_fastq(text, n, final_batch, recs, cap, consumed, err, err_rec);
    return parse_fasta(text, n, final_batch, recs, cap, consumed, err, err_rec);
}

size_t orc_seq(const uint8_t* text, const orc_rec* r, int fmt, uint8_t* out, size_t cap) {
    (void)fmt;
    size_t len = 0;
    for (uint64_t q = r->seq_off; q < r->seq_end; ++q) {
        uint8_t c = text[q];
        if (c == '\n') continue;
        if (c == '\r' && q + 1 < r->seq_end && text[q + 1] == '\n') continue; /* CRLF inside */
        if (len < cap && out) out[len] = c;
        ++len;
    }
    return len;
}

int orc_assign(const uint8_t* text, const orc_rec* recs, size_t n_rec, int fmt,
               const uint8_t* barcodes, const uint32_t* bc_lens, uint32_t n_barcodes,
               uint32_t stride, uint8_t mismatch, uint16_t* assign, size_t* err_rec) {
    uint32_t bc_len = bc_lens[0]; /* "assume all barcodes have same length", demux.rs:30-34 */
    uint8_t* pre = (uint8_t*)malloc(bc_len ? bc_len : 1);
    for (size_t i = 0; i < n_rec; ++i) {
        size_t len = orc_seq(text, &recs[i], fmt, pre, bc_len);
        if (len < bc_len) { *err_rec = i; free(pre); return ORC_E_SHORT_READ; } /* slice panic */
        uint32_t hit = n_barcodes;
        for (uint32_t b = 0; b < n_barcodes; ++b) { /* iter().find(): first hit wins */
            if (orc_bc_cmp(barcodes + (size_t)b * stride, bc_lens[b], pre, bc_len, mismatch)) {
                hit = b;
                break;
            }
        }
        assign[i] = (uint16_t)hit;
    }
    free(pre);
    return ORC_OK;
}

void orc_partition(const uint16_t* assign, size_t n_rec, uint32_t nbins, uint32_t* bin_count,
                   uint32_t* bin_start, uint32_t* order) {
    memset(bin_count, 0, sizeof(uint32_t) * nbins);
    for (size_t i = 0; i < n_rec; ++i) bin_count[assign[i]]++;
    uint32_t acc = 0;
    for (uint32_t b = 0; b < nbins; ++b) { bin_start[b] = acc; acc += bin_count[b]; }
    bin_start[nbins] = acc;
    uint32_t* cur = (uint32_t*)malloc(sizeof(uint32_t) * nbins);
    memcpy(cur, bin_start, sizeof(uint32_t) * nbins);
    for (size_t i = 0; i < n_rec; ++i) order[cur[assign[i]]++] = (uint32_t)i;
    free(cur);
}

size_t orc_serialize(const uint8_t* text, const orc_rec* r, int fmt, uint8_t* out) {
    size_t w = 0;
#define PUT(c) do { if (out) out[w] = (uint8_t)(c); ++w; } while (0)
    PUT(fmt == ORC_FMT_FASTQ ? '@' : '>');
    for (uint64_t q = 0; q < r->id_len; ++q) PUT(text[r->id_off + q]);
    PUT('\n');
    w += orc_seq(text, r, fmt, out ? out + w : NULL, (size_t)-1);
    PUT('\n');
    if (fmt == ORC_FMT_FASTQ) {
        PUT('+');
        PUT('\n');
        for (uint64_t q = 0; q < r->qual_len; ++q) PUT(text[r->qual_off + q]);
        PUT('\n');
    }
#undef PUT
    return w;
}

uint64_t orc_fnv1a(const uint8_t* p, size_t n, uint64_t h) {
    if (h == 0) h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return h;
}

#define SINK_CAP (64u * 1024u)

long orc_demux_stream(const uint8_t* text, size_t n, const uint8_t* barcodes,
                      const uint32_t* bc_lens, uint32_t n_barcodes, uint32_t stride,
                      uint8_t mismatch, uint64_t* bin_count, uint64_t* sink_bytes,
                      uint64_t* sink_hash, int keep_bytes) {
    (void)keep_bytes;
    uint32_t nbins = n_barcodes + 1;
    uint32_t bc_len = bc_lens[0];
    static const uint8_t sentinel[3] = {'X', 'X', 'X'}; /* main.rs:115,158: the "XXX" key */
    uint8_t** sink = (uint8_t**)malloc(sizeof(uint8_t*) * nbins);
    uint32_t* fill = (uint32_t*)calloc(nbins, sizeof(uint32_t));
    for (uint32_t b = 0; b < nbins; ++b) sink[b] = (uint8_t*)malloc(SINK_CAP);
    memset(bin_count, 0, sizeof(uint64_t) * nbins);
    memset(sink_bytes, 0, sizeof(uint64_t) * nbins);
    if (sink_hash) memset(sink_hash, 0, sizeof(uint64_t) * nbins);
    uint8_t* pre = (uint8_t*)malloc(bc_len + 1);
    uint8_t* tmp = NULL;
    size_t tmp_cap = 0;
    long done = 0;
    size_t pos = 0;
    int fmt = 0;
    long rc = 0;
    while (pos < n) {
        orc_rec r;
        size_t consumed = 0, err_rec = 0;
        int err = 0;
        /* one record at a time, like reader.next() (demux.rs:49): parse a window that is
         * guaranteed to hold the next record, capacity 1 */
        size_t win = n - pos;
        long got = orc_parse(text + pos, win, 1, &r, 1, &fmt, &consumed, &err, &err_rec);
        if (got < 0 && err != ORC_E_CAPACITY) { rc = -1; break; }
        if (got == 0 && err == ORC_OK) break; /* blank tail */
        /* err == CAPACITY means "more than one record available": r holds the first */
        size_t len = orc_seq(text + pos, &r, fmt, pre, bc_len);
        if (len < bc_len) { rc = -1; break; }
        uint32_t hit = nbins;
        for (uint32_t b = 0; b < n_barcodes; ++b)
            if (orc_bc_cmp(barcodes + (size_t)b * stride, bc_lens[b], pre, bc_len, mismatch)) { hit = b; break; }
        if (hit == nbins) {
            (void)orc_bc_cmp(sentinel, 3, pre, bc_len, mismatch); /* the sentinel is a candidate too */
            hit = n_barcodes;
        }
        size_t need = (size_t)(r.rec_end - r.rec_off) + 8;
        if (need > tmp_cap) { tmp_cap = need * 2; tmp = (uint8_t*)realloc(tmp, tmp_cap); }
        size_t w = orc_serialize(text + pos, &r, fmt, tmp);
        if (sink_hash) sink_hash[hit] = orc_fnv1a(tmp, w, sink_hash[hit]);
        /* buffered append (stands in for write_seqs' 5 / 9 write_all calls) */
        size_t off = 0;
        while (off < w) {
            size_t room = SINK_CAP - fill[hit];
            size_t c = w - off < room ? w - off : room;
            memcpy(sink[hit] + fill[hit], tmp + off, c);
            fill[hit] += (uint32_t)c;
            off += c;
            if (fill[hit] == SINK_CAP) fill[hit] = 0; /* "flush" */
        }
        bin_count[hit]++;
        sink_bytes[hit] += w;
        pos += (size_t)r.rec_end;
        ++done;
    }
    for (uint32_t b = 0; b < nbins; ++b) free(sink[b]);
    free(sink); free(fill); free(pre); free(tmp);
    return rc < 0 ? -1 : done;
}
