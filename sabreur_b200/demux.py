"""Host-side mirror of the reference's demux interface over the C ABI (include/sabreur_gpu.h).

  Demux                  one GPU context: barcode table (file order), mismatch budget, pinned slots
  Demux.stream(chunks)   push a decompressed byte stream through the GPU, yields one Batch per submit
  se_demux / pe_demux    the reference's entry points (src/demux.rs:13-20, 71-79): same arguments in
                         spirit (paths, output format/level, barcode -> output files, mismatch, counters)
                         and the same return values ((nb_records, is_unk_empty) / (nb_records, "truefalse"))

All compute happens in libsabreur_gpu.so; this file only moves bytes and writes files.  Nothing here
falls back to a CPU implementation: without the library or a GPU every call raises.
"""
from __future__ import annotations

import bz2
import ctypes as C
import gzip
import lzma
import os
from dataclasses import dataclass
from typing import Iterable, Iterator

import numpy as np

from . import _ffi

FMT_AUTO, FMT_FASTA, FMT_FASTQ = _ffi.FMT_AUTO, _ffi.FMT_FASTA, _ffi.FMT_FASTQ


class SabreurGpuError(RuntimeError):
    pass


class ParseError(SabreurGpuError):
    """the batch holds a record the reference's parser would refuse (or a read shorter than bc_len)"""

    def __init__(self, flags: int, record: int):
        names = [n for bit, n in ((1, "invalid start"), (2, "invalid separator"), (4, "unequal lengths"),
                                  (8, "read shorter than barcode"), (16, "truncated record")) if flags & bit]
        super().__init__(f"record {record}: " + ", ".join(names))
        self.flags, self.record = flags, record


@dataclass
class Batch:
    seq: int
    text: np.ndarray          # u8 view of the batch bytes (valid until the slot is reused)
    consumed: int
    n_records: int
    status: int
    first_bad_record: int
    first_bad_flags: int
    fmt: int
    rec_off: np.ndarray       # u32 [n+1]
    assign: np.ndarray        # u16 [n]
    bin_count: np.ndarray     # u32 [n_bins]
    bin_start: np.ndarray     # u32 [n_bins+1]
    order: np.ndarray         # u32 [n]
    rec_key: np.ndarray | None
    scan_path: int = 0        # 1: region-parallel scan, 0: look-back scan
    record_base: int = 0      # index of the batch's first record in the whole stream


def _as_np(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    ct = {np.uint32: C.c_uint32, np.uint16: C.c_uint16, np.uint64: C.c_uint64, np.uint8: C.c_uint8}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))


class Demux:
    def __init__(self, barcodes: list[bytes], mismatch: int = 0, fmt: int = FMT_AUTO, device: int = 0,
                 batch_bytes: int = 64 << 20, n_slots: int = 2, max_records: int = 0, flags: int = 0):
        self.lib = _ffi.load()
        bcs = [bytes(b) for b in barcodes]
        if not bcs or not bcs[0]:
            raise ValueError("Barcode data is empty")  # src/demux.rs:31-33
        k = len(bcs[0])  # bc_len comes from the first row (src/demux.rs:34, refinement H2)
        tab = np.zeros((len(bcs), k), dtype=np.uint8)
        lens = np.zeros(len(bcs), dtype=np.uint32)
        for i, b in enumerate(bcs):
            b = b[:k]  # zip() against seq[..bc_len] never looks past bc_len
            tab[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
            lens[i] = len(b)
        self._tab, self._lens = tab, lens
        self.n_barcodes, self.bc_len, self.mismatch = len(bcs), k, mismatch
        self.n_bins = len(bcs) + 1
        self.batch_bytes, self.n_slots, self.device = batch_bytes, n_slots, device
        cfg = _ffi.SbrConfig(device=device, fmt=fmt, n_barcodes=len(bcs), bc_len=k, barcodes=tab.ctypes.data,
                             bc_lens=lens.ctypes.data, mismatch=mismatch, n_slots=n_slots, batch_bytes=batch_bytes,
                             max_records=max_records, flags=flags)
        h = C.c_void_p()
        rc = self.lib.sbr_create(C.byref(h), C.byref(cfg))
        if rc != 0:
            raise SabreurGpuError(f"sbr_create failed ({rc}): {self.lib.sbr_last_error(None).decode()}")
        self.h = h
        self._slot_ptr: dict[int, int] = {}

    # -- plumbing -----------------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise SabreurGpuError(f"{what} failed ({rc}): {self.lib.sbr_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.sbr_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def slot_buffer(self, slot: int) -> np.ndarray:
        if slot not in self._slot_ptr:
            p, cap = C.c_void_p(), C.c_uint64()
            self._check(self.lib.sbr_slot_buffer(self.h, slot, C.byref(p), C.byref(cap)), "sbr_slot_buffer")
            self._slot_ptr[slot] = p.value
        return _as_np(self._slot_ptr[slot], self.batch_bytes, np.uint8)

    def submit(self, slot: int, offset: int, nbytes: int, seq: int, final: bool):
        self._check(self.lib.sbr_submit(self.h, slot, offset, nbytes, seq, int(final)), "sbr_submit")

    def wait(self, slot: int, copy: bool = True) -> Batch:
        r = _ffi.SbrResult()
        self._check(self.lib.sbr_wait(self.h, slot, C.byref(r)), "sbr_wait")
        n = r.n_records
        f = (lambda a: a.copy()) if copy else (lambda a: a)
        return Batch(seq=r.batch_seq, text=np.zeros(0, np.uint8), consumed=r.consumed, n_records=n,
                     status=r.status_flags, first_bad_record=r.first_bad_record, first_bad_flags=r.first_bad_flags,
                     fmt=r.fmt, rec_off=f(_as_np(r.rec_off, n + 1, np.uint32)), assign=f(_as_np(r.assign, n, np.uint16)),
                     bin_count=f(_as_np(r.bin_count, r.n_bins, np.uint32)),
                     bin_start=f(_as_np(r.bin_start, r.n_bins + 1, np.uint32)), order=f(_as_np(r.order, n, np.uint32)),
                     rec_key=f(_as_np(r.rec_key, n, np.uint64)) if r.rec_key else None, scan_path=r.scan_path)

    def demux_device(self, slot: int, d_ptr: int, nbytes: int, final: bool = True, stream: int = 0):
        self._check(self.lib.sbr_demux_device(self.h, slot, d_ptr, nbytes, int(final), stream), "sbr_demux_device")

    def join(self, stream: int = 0):
        """CFG_OVERLAP: `stream` waits for the match + partition work outstanding on the slots' own streams"""
        self._check(self.lib.sbr_device_join(self.h, stream), "sbr_device_join")

    def scan_stats(self, reset: bool = False) -> tuple[int, int]:
        """(batches the region-parallel scan finished, batches it handed over to the exact scan)"""
        out = (C.c_uint64 * 2)()
        self._check(self.lib.sbr_scan_stats(self.h, C.byref(out), int(reset)), "sbr_scan_stats")
        return int(out[0]), int(out[1])

    def scan_geometry(self, fmt: int) -> tuple[int, int]:
        """(tile bytes, regions) of the region-parallel scan: batches of about regions * tile_bytes * k bytes load every
        region evenly"""
        t, r = C.c_uint32(), C.c_uint32()
        self._check(self.lib.sbr_scan_geometry(self.h, fmt, C.byref(t), C.byref(r)), "sbr_scan_geometry")
        return int(t.value), int(r.value)

    def lut(self) -> np.ndarray | None:
        n = self.lib.sbr_lut_entries(self.h)
        if n == 0:
            return None
        out = np.zeros(n, dtype=np.uint16)
        self._check(self.lib.sbr_lut_copy(self.h, out.ctypes.data, n), "sbr_lut_copy")
        return out

    def timing(self):
        ms = (C.c_double * 3)()
        b, l = C.c_uint64(), C.c_uint64()
        self._check(self.lib.sbr_timing_collect(self.h, C.byref(ms), C.byref(b), C.byref(l)), "sbr_timing_collect")
        return list(ms), b.value, l.value

    # -- streaming ------------------------------------------------------------------------------------
    def stream(self, chunks: Iterable[bytes], reserve: int | None = None) -> Iterator[Batch]:
        """Feeds a byte stream through the GPU in record-aligned batches.

        The unconsumed tail of batch i (an incomplete record) is prepended to batch i+1: room for it is
        reserved at the front of every slot, so freshly read bytes are never moved.  Batches are therefore
        chained (i+1 is submitted once i's `consumed` is known); the exactness of record boundaries comes
        from the GPU scan alone, never from a host-side guess.  What does overlap: while the GPU works on batch i
        the host already reads (decompresses) the fresh bytes of batch i+1 into the next slot, behind the reserved room.
        Yields Batch objects whose .text views the slot memory: use it before asking for the next batch.
        """
        cap = self.batch_bytes
        reserve = cap // 4 if reserve is None else reserve
        it = iter(chunks)
        pending = b""      # bytes read from `chunks` but not yet placed in a slot
        eof = False

        def fill(slot: int) -> tuple[int, bool]:
            """fresh bytes into the slot behind the reserved room -> (bytes placed, the stream ends with them)"""
            nonlocal pending, eof
            buf = self.slot_buffer(slot)
            filled = 0
            while reserve + filled < cap and (pending or not eof):
                if not pending:
                    try:
                        pending = next(it)
                    except StopIteration:
                        eof = True
                        break
                    if not pending:
                        continue
                take = min(len(pending), cap - reserve - filled)
                buf[reserve + filled: reserve + filled + take] = np.frombuffer(pending[:take], dtype=np.uint8)
                pending = pending[take:]
                filled += take
            if not pending and not eof:  # learn about EOF before deciding whether this batch is final
                try:
                    pending = next(it)
                except StopIteration:
                    eof = True
            return filled, eof and not pending

        tail = np.zeros(0, dtype=np.uint8)
        seq = 0
        base = 0
        slot = 0
        filled, final = fill(slot)
        while True:
            buf = self.slot_buffer(slot)
            if tail.size <= reserve:
                off = reserve - tail.size
                buf[off:reserve] = tail
            else:  # a tail larger than the reserved room (a record of more than a quarter batch): slide the fresh bytes back
                over = max(0, tail.size + filled - cap)
                if over:  # ... and hand what no longer fits back to the reader
                    pending = buf[reserve + filled - over: reserve + filled].tobytes() + pending
                    filled -= over
                    final = False
                buf[tail.size: tail.size + filled] = buf[reserve: reserve + filled].copy()
                buf[:tail.size] = tail
                off = 0
            nbytes = tail.size + filled
            if nbytes == 0:
                if seq == 0:
                    raise ParseError(_ffi.ST_INVALID_START, 0)  # empty input: needletail refuses it
                return
            self.submit(slot, off, nbytes, seq, final)
            nxt = (slot + 1) % self.n_slots
            prefilled = None
            if not final and nxt != slot:
                prefilled = fill(nxt)  # the host reads ahead while the GPU scans
            b = self.wait(slot)
            b.text = buf[off: off + nbytes]
            b.record_base = base
            yield b
            if b.status & _ffi.ST_ERROR_MASK:
                return
            base += b.n_records
            tail = buf[off + b.consumed: off + nbytes].copy()
            if b.consumed == 0 and not final:
                raise SabreurGpuError("no progress: a single record is larger than a batch")
            seq += 1
            if final and not ((b.status & _ffi.ST_REC_OVERFLOW) and tail.size):
                return
            slot = nxt
            filled, final = prefilled if prefilled is not None else fill(slot)


def demux_bytes(data: bytes, barcodes: list[bytes], mismatch: int, fmt: int = FMT_AUTO, batch_bytes: int = 8 << 20,
                chunk: int = 1 << 20, **kw):
    """whole-buffer convenience: -> dict(assign, bin_count, order(global, grouped by bin), n_records, status...)"""
    with Demux(barcodes, mismatch, fmt=fmt, batch_bytes=batch_bytes, **kw) as dm:
        asg, counts = [], np.zeros(dm.n_bins, dtype=np.uint64)
        per_bin = [[] for _ in range(dm.n_bins)]
        spans = []
        status, bad = 0, None
        pos = 0
        fmt_seen = 0
        paths = []
        for b in dm.stream(data[i:i + chunk] for i in range(0, len(data), chunk)):
            status |= b.status
            fmt_seen = b.fmt or fmt_seen
            paths.append(b.scan_path)
            if b.status & _ffi.ST_ERROR_MASK:
                bad = (b.record_base + b.first_bad_record, b.first_bad_flags)
                # records before the first bad one were still demultiplexed by the reference
                nok = b.first_bad_record
            else:
                nok = b.n_records
            asg.append(b.assign[:nok])
            counts += np.bincount(b.assign[:nok], minlength=dm.n_bins).astype(np.uint64)
            for bin_ in range(dm.n_bins):
                seg = b.order[b.bin_start[bin_]: b.bin_start[bin_ + 1]]
                per_bin[bin_].append(seg[seg < nok].astype(np.uint64) + b.record_base)
            spans.append(np.stack([b.rec_off[:nok].astype(np.uint64) + pos, b.rec_off[1:nok + 1].astype(np.uint64) + pos], 1))
            pos += b.consumed
        return dict(assign=np.concatenate(asg) if asg else np.zeros(0, np.uint16), bin_count=counts,
                    per_bin=[np.concatenate(x) if x else np.zeros(0, np.uint64) for x in per_bin],
                    spans=np.concatenate(spans) if spans else np.zeros((0, 2), np.uint64), status=status, bad=bad,
                    fmt=fmt_seen, scan_paths=paths)


# ---------------------------------------------------------------------------------------------------
# canonical re-serialisation (src/utils.rs:133-158): only for records flagged non-canonical
# ---------------------------------------------------------------------------------------------------
def _canonical(rec: bytes, fmt: int) -> bytes:
    lines = rec.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    lines = [ln[:-1] if ln.endswith(b"\r") else ln for ln in lines]
    if fmt == FMT_FASTQ:
        return b"@" + lines[0][1:] + b"\n" + lines[1] + b"\n+\n" + lines[3] + b"\n"
    return b">" + lines[0][1:] + b"\n" + b"".join(lines[1:]) + b"\n"


_OPENERS = {"": open, ".gz": gzip.open, ".bz2": bz2.open, ".xz": lzma.open}
_MAGIC = ((b"\x1f\x8b", ".gz"), (b"BZ", ".bz2"), (b"\xfd7zXZ", ".xz"), (b"\x28\xb5\x2f\xfd", ".zst"))


def which_format(path: str) -> str:
    """niffler::sniff as used by src/utils.rs:125-130 -> extension string ('' = uncompressed)"""
    with open(path, "rb") as fh:
        head = fh.read(5)
    if len(head) < 5:
        raise SabreurGpuError("file too short")  # niffler: FileTooShort
    for magic, ext in _MAGIC:
        if head.startswith(magic):
            return ext
    return ""


def _read_chunks(path: str, chunk: int = 4 << 20) -> Iterator[bytes]:
    ext = which_format(path)
    if ext == ".zst":
        raise SabreurGpuError("zstd input needs the C++ host (no zstd module in this Python)")
    with _OPENERS[ext](path, "rb") as fh:
        while True:
            b = fh.read(chunk)
            if not b:
                return
            yield b


def _member(data: bytes, fmt: str, level: int) -> bytes:
    """one compressed member holding `data` (members concatenate into a valid file; the reference writes one member
    per record through niffler::get_writer, src/utils.rs:139 -- same decompressed content)"""
    if fmt == "":
        return data
    if fmt == ".gz":
        return gzip.compress(data, compresslevel=level)
    if fmt == ".bz2":
        return bz2.compress(data, compresslevel=level)
    if fmt == ".xz":
        return lzma.compress(data, preset=level)
    raise SabreurGpuError("zstd output needs the C++ host (no zstd module in this Python)")


def _demux_one(path: str, writers: list, unknown_writer, barcodes: list[bytes], mismatch: int,
               nb_records: dict, strict: bool, fmt: str = "", level: int = 1, **kw) -> bool:
    """one pass of the hot loop (src/demux.rs:49-66 / 101-110 / 114-123). -> is_unk_empty"""
    unk_empty = True
    with Demux(barcodes, mismatch, **kw) as dm:
        for b in dm.stream(_read_chunks(path)):
            nok = b.first_bad_record if (b.status & _ffi.ST_ERROR_MASK) else b.n_records
            text = b.text
            noncanon = b.rec_key is not None
            for bin_ in range(dm.n_bins):
                seg = b.order[b.bin_start[bin_]: b.bin_start[bin_ + 1]]
                seg = seg[seg < nok]
                if seg.size == 0:
                    continue
                w = unknown_writer if bin_ == dm.n_barcodes else writers[bin_]
                if bin_ == dm.n_barcodes:
                    unk_empty = False
                else:
                    nb_records[barcodes[bin_]] = nb_records.get(barcodes[bin_], 0) + int(seg.size)
                parts = []
                for i in seg:
                    rec = text[b.rec_off[i]: b.rec_off[i + 1]].tobytes()
                    if noncanon and (b.fmt == FMT_FASTA or (int(b.rec_key[i]) & _ffi.KEY_NONCANON)):
                        rec = _canonical(rec, b.fmt)
                    parts.append(rec)
                w.write(_member(b"".join(parts), fmt, level))
            if b.status & _ffi.ST_ERROR_MASK:
                if b.record_base + b.first_bad_record == 0 and (b.first_bad_flags & _ffi.ST_INVALID_START):
                    # parse_fastx_reader(..)? fails before the loop starts (demux.rs:23, 85-86): SE and PE alike
                    raise ParseError(b.first_bad_flags, 0)
                if b.first_bad_flags & _ffi.ST_SHORT_READ:
                    raise ParseError(b.first_bad_flags, b.record_base + b.first_bad_record)  # slice panic, demux.rs:54
                if strict:
                    raise ParseError(b.first_bad_flags, b.record_base + b.first_bad_record)  # `record?`, demux.rs:50
                break  # `while let Some(Ok(record))` stops silently, demux.rs:101,114
    return unk_empty


_LEVELS = {1, 2, 3, 4, 5, 6, 7, 8, 9}


def se_demux(file: str, format: str, level: int, barcode_data: dict, mismatch: int, nb_records: dict, **kw):
    """src/demux.rs:13-67, same argument order.  `format`: "" (niffler Format::No: keep the input's compression),
    ".gz", ".bz2", ".xz" (".zst" needs the C++ host); `level` 1-9 (utils.rs:87-100: anything else -> 1).
    barcode_data: {barcode bytes: [writable binary file]} in barcode-file order, with the b"XXX" entry holding the
    unknown writer (src/main.rs:152-158).  -> (nb_records, is_unk_empty)"""
    if not barcode_data:
        raise ValueError("Barcode data is empty")
    if b"XXX" not in barcode_data:
        raise ValueError("Missing 'XXX' fallback barcode entry in barcode_data")
    if format == "":
        format = which_format(file)  # demux.rs:26-29
    level = level if level in _LEVELS else 1
    bcs = [k for k in barcode_data if k != b"XXX"]
    writers = [barcode_data[k][0] for k in bcs]
    unk = _demux_one(file, writers, barcode_data[b"XXX"][0], bcs, mismatch, nb_records, strict=True, fmt=format, level=level,
                     **kw)
    return nb_records, unk


def pe_demux(forward: str, reverse: str, format: str, level: int, barcode_data: dict, mismatch: int, nb_records: dict,
             **kw):
    """src/demux.rs:71-128: two independent passes, forward then reverse; both mates are written with the
    forward-derived compression unless `format` is given (demux.rs:81,95-97).  -> (nb_records, "truetrue"...)"""
    compression = which_format(forward)
    which_format(reverse)
    if format != "":
        compression = format
    for path in (forward, reverse):  # both parsers are built before the first record is processed (demux.rs:85-86)
        first = next(_read_chunks(path, 1), b"")
        if first[:1] not in (b">", b"@"):
            raise ParseError(_ffi.ST_INVALID_START, 0)
    level = level if level in _LEVELS else 1
    bcs = [k for k in barcode_data if k != b"XXX"]
    unk = barcode_data[b"XXX"]
    u1 = _demux_one(forward, [barcode_data[k][0] for k in bcs], unk[0], bcs, mismatch, nb_records, strict=False,
                    fmt=compression, level=level, **kw)
    u2 = _demux_one(reverse, [barcode_data[k][1] for k in bcs], unk[1], bcs, mismatch, nb_records, strict=False,
                    fmt=compression, level=level, **kw)
    return nb_records, f"{str(u1).lower()}{str(u2).lower()}"
