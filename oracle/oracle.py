"""ctypes loader for the C oracle + an independent pure-Python twin for tiny cases.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the CPU-baseline legs
of bench.py; never by the product package (sabreur_b200/).

Two restatements of the same reference code (C in sabreur_oracle.c, Python here) are kept so that
they can be checked against each other: the reference binary itself cannot be built in this
environment (no Rust toolchain), see DESIGN.md "Oracle".
Reference: /root/reference/src/utils.rs:115-123 (bc_cmp), src/demux.rs:49-66, 101-123 (match loop),
src/utils.rs:133-158 (serialisation).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FMT_FASTA, FMT_FASTQ = 1, 2
E_OK, E_EMPTY, E_INVALID_START, E_INVALID_SEP, E_UNEQUAL_LEN, E_TRUNCATED, E_SHORT_READ, E_CAPACITY = range(8)
REC_NONCANONICAL = 1


class OrcRec(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("rec_off", "rec_end", "id_off", "id_len", "seq_off", "seq_end", "qual_off", "qual_len")] + \
               [("flags", C.c_uint32), ("pad", C.c_uint32)]


REC_DTYPE = np.dtype([(n, "<u8") for n in
                      ("rec_off", "rec_end", "id_off", "id_len", "seq_off", "seq_end", "qual_off", "qual_len")] +
                     [("flags", "<u4"), ("pad", "<u4")])


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("sabreur_oracle.c", "synth_ref.c")]
    hdr = os.path.join(_HERE, "sabreur_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs + [hdr]):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", so] + srcs)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p = C.c_void_p
        L.orc_bc_cmp.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint8]
        L.orc_bc_cmp.restype = C.c_int
        L.orc_parse.argtypes = [u8p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int),
                                C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        L.orc_parse.restype = C.c_long
        L.orc_seq.argtypes = [u8p, C.c_void_p, C.c_int, u8p, C.c_size_t]
        L.orc_seq.restype = C.c_size_t
        L.orc_assign.argtypes = [u8p, C.c_void_p, C.c_size_t, C.c_int, u8p, C.c_void_p, C.c_uint32, C.c_uint32,
                                 C.c_uint8, C.c_void_p, C.POINTER(C.c_size_t)]
        L.orc_assign.restype = C.c_int
        L.orc_partition.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_partition.restype = None
        L.orc_serialize.argtypes = [u8p, C.c_void_p, C.c_int, u8p]
        L.orc_serialize.restype = C.c_size_t
        L.orc_demux_stream.argtypes = [u8p, C.c_size_t, u8p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint8,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_demux_stream.restype = C.c_long
        L.orc_fnv1a.argtypes = [u8p, C.c_size_t, C.c_uint64]
        L.orc_fnv1a.restype = C.c_uint64
        L.orc_synth_table.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, u8p]
        L.orc_synth_table.restype = C.c_int
        L.orc_synth_record_bytes.argtypes = [C.c_int, C.c_uint32]
        L.orc_synth_record_bytes.restype = C.c_uint32
        L.orc_synth_reads.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, u8p, C.c_uint32, C.c_uint32,
                                      C.c_uint64, C.c_uint64, u8p]
        L.orc_synth_reads.restype = C.c_int
        _LIB = L
    return _LIB


def synth_table(seed: int, n_barcodes: int, bc_len: int, min_dist: int) -> np.ndarray:
    """greedy min-distance barcode table (oracle/synth_ref.c)"""
    out = np.zeros((n_barcodes, bc_len), dtype=np.uint8)
    if lib().orc_synth_table(seed, n_barcodes, bc_len, min_dist, out.ctypes.data) != 0:
        raise RuntimeError("orc_synth_table failed")
    return out


def synth_reads(fmt: int, read_len: int, mate: int, seed: int, table: np.ndarray, first: int, n: int,
                threads: int = 1) -> np.ndarray:
    """records [first, first + n) of the synthetic stream (oracle/synth_ref.c); `threads` > 1 fills
    disjoint ranges in parallel (ctypes releases the GIL)"""
    L = lib()
    rec = L.orc_synth_record_bytes(int(fmt == FMT_FASTQ), read_len)
    out = np.empty(n * rec, dtype=np.uint8)
    tab = np.ascontiguousarray(table, dtype=np.uint8)

    def fill(lo, hi):
        rc = L.orc_synth_reads(int(fmt == FMT_FASTQ), read_len, mate, seed, tab.ctypes.data, tab.shape[0], tab.shape[1],
                               first + lo, hi - lo, out.ctypes.data + lo * rec)
        if rc != 0:
            raise RuntimeError("orc_synth_reads failed")

    if threads <= 1 or n < 4 * threads:
        fill(0, n)
    else:
        import threading
        per = (n + threads - 1) // threads
        ths = [threading.Thread(target=fill, args=(i * per, min(n, (i + 1) * per))) for i in range(threads) if i * per < n]
        [t.start() for t in ths]
        [t.join() for t in ths]
    return out


class OracleError(Exception):
    def __init__(self, code: int, record: int, consumed: int = 0):
        super().__init__(f"oracle parse error code={code} record={record}")
        self.code, self.record, self.consumed = code, record, consumed


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def bc_cmp(bc: bytes, seq: bytes, mismatch: int) -> bool:
    return bool(lib().orc_bc_cmp(bytes(bc), len(bc), bytes(seq), len(seq), mismatch))


def pack_table(barcodes) -> tuple[np.ndarray, np.ndarray, int]:
    """rows of raw ASCII -> (n x stride u8 matrix, lens u32, stride)"""
    bcs = [bytes(b) for b in barcodes]
    stride = max(1, max(len(b) for b in bcs))
    tab = np.zeros((len(bcs), stride), dtype=np.uint8)
    for i, b in enumerate(bcs):
        tab[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    lens = np.array([len(b) for b in bcs], dtype=np.uint32)
    return tab, lens, stride


def parse(text, final: bool = True, fmt: int = 0, cap: int | None = None):
    """-> (recs structured array, fmt, consumed).  Raises OracleError like the reference's `record?`."""
    t = _u8(text)
    n = t.size
    if cap is None:
        cap = n // 2 + 2
    recs = np.zeros(cap, dtype=REC_DTYPE)
    f = C.c_int(fmt)
    consumed, err, err_rec = C.c_size_t(0), C.c_int(0), C.c_size_t(0)
    got = lib().orc_parse(t.ctypes.data, n, int(final), recs.ctypes.data, cap, C.byref(f), C.byref(consumed),
                          C.byref(err), C.byref(err_rec))
    if got < 0:
        raise OracleError(err.value, err_rec.value, consumed.value)
    return recs[:got].copy(), f.value, consumed.value


def assign(text, recs, fmt, barcodes, mismatch: int) -> np.ndarray:
    t = _u8(text)
    tab, lens, stride = pack_table(barcodes)
    out = np.zeros(len(recs), dtype=np.uint16)
    recs = np.ascontiguousarray(recs)
    err_rec = C.c_size_t(0)
    rc = lib().orc_assign(t.ctypes.data, recs.ctypes.data, len(recs), fmt, tab.ctypes.data, lens.ctypes.data,
                          len(lens), stride, mismatch, out.ctypes.data, C.byref(err_rec))
    if rc != 0:
        raise OracleError(rc, err_rec.value)
    return out


def partition(assign_arr: np.ndarray, nbins: int):
    a = np.ascontiguousarray(assign_arr, dtype=np.uint16)
    cnt = np.zeros(nbins, dtype=np.uint32)
    start = np.zeros(nbins + 1, dtype=np.uint32)
    order = np.zeros(a.size, dtype=np.uint32)
    lib().orc_partition(a.ctypes.data, a.size, nbins, cnt.ctypes.data, start.ctypes.data, order.ctypes.data)
    return cnt, start, order


def serialize(text, rec, fmt) -> bytes:
    t = _u8(text)
    r = np.ascontiguousarray(np.array([rec], dtype=REC_DTYPE))
    n = lib().orc_serialize(t.ctypes.data, r.ctypes.data, fmt, None)
    out = np.zeros(n, dtype=np.uint8)
    lib().orc_serialize(t.ctypes.data, r.ctypes.data, fmt, out.ctypes.data)
    return out.tobytes()


def seq(text, rec, fmt) -> bytes:
    t = _u8(text)
    r = np.ascontiguousarray(np.array([rec], dtype=REC_DTYPE))
    n = lib().orc_seq(t.ctypes.data, r.ctypes.data, fmt, None, 0)
    out = np.zeros(max(n, 1), dtype=np.uint8)
    lib().orc_seq(t.ctypes.data, r.ctypes.data, fmt, out.ctypes.data, n)
    return out[:n].tobytes()


def demux(text, barcodes, mismatch: int, final: bool = True, fmt: int = 0):
    """whole-buffer demux -> dict(recs, fmt, consumed, assign, bin_count, bin_start, order)"""
    recs, f, consumed = parse(text, final=final, fmt=fmt)
    a = assign(text, recs, f, barcodes, mismatch)
    cnt, start, order = partition(a, len(barcodes) + 1)
    return dict(recs=recs, fmt=f, consumed=consumed, assign=a, bin_count=cnt, bin_start=start, order=order)


def demux_outputs(text, barcodes, mismatch: int) -> list[bytes]:
    """canonical per-bin output streams (bin n_barcodes = unknown), src/utils.rs:133-158"""
    d = demux(text, barcodes, mismatch)
    outs = [bytearray() for _ in range(len(barcodes) + 1)]
    for i, r in enumerate(d["recs"]):
        outs[int(d["assign"][i])] += serialize(text, r, d["fmt"])
    return [bytes(o) for o in outs]


def demux_stream(text, barcodes, mismatch: int, want_hash: bool = False):
    """the reference-faithful one-record-at-a-time loop (timed CPU baseline)"""
    t = _u8(text)
    tab, lens, stride = pack_table(barcodes)
    nb = len(lens) + 1
    cnt = np.zeros(nb, dtype=np.uint64)
    sb = np.zeros(nb, dtype=np.uint64)
    sh = np.zeros(nb, dtype=np.uint64)
    got = lib().orc_demux_stream(t.ctypes.data, t.size, tab.ctypes.data, lens.ctypes.data, len(lens), stride,
                                 mismatch, cnt.ctypes.data, sb.ctypes.data, sh.ctypes.data if want_hash else None, 0)
    if got < 0:
        raise OracleError(-1, 0)
    return got, cnt, sb, sh


def fnv1a(data: bytes, h: int = 0) -> int:
    t = _u8(data)
    return int(lib().orc_fnv1a(t.ctypes.data, t.size, h))


# ---------------------------------------------------------------------------------------------
# Independent pure-Python twin (tiny inputs only).  Written from the reference source, not from
# the C file: it exists to cross-check the C oracle.
# ---------------------------------------------------------------------------------------------

def py_bc_cmp(bc: bytes, seq_: bytes, mismatch: int) -> bool:
    # src/utils.rs:115-123: zip truncates; sum::<u8>() wraps in the release profile
    return (sum(1 for a, b in zip(bc, seq_) if a != b) & 0xFF) <= mismatch


def _trim_cr(b: bytes) -> bytes:
    return b[:-1] if b.endswith(b"\r") else b


def py_records(text: bytes):
    """-> (fmt, [(id, seq, qual|None)]) for a complete (final) input; raises ValueError on bad input"""
    if not text:
        raise ValueError("empty")
    if text[:1] == b"@":
        lines = text.split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()  # the split artefact after a final newline
        recs = []
        i = 0
        while i < len(lines):
            rest = lines[i:]
            if len(rest) < 4:
                if all(_trim_cr(x) == b"" for x in rest):
                    break
                if len(rest) < 3 or not text.endswith(b"\n"):
                    raise ValueError("truncated")
                rest = rest + [b""]  # three newline-terminated lines at EOF: empty quality line
            h, s, p, q = (_trim_cr(x) for x in rest[:4])
            if not h.startswith(b"@"):
                raise ValueError("invalid start")
            if not p.startswith(b"+"):
                raise ValueError("invalid separator")
            if len(s) != len(q):
                raise ValueError("unequal lengths")
            recs.append((h[1:], s, q))
            i += 4
        return FMT_FASTQ, recs
    if text[:1] == b">":
        recs = []
        for chunk in (b"\n" + text).split(b"\n>")[1:]:
            parts = chunk.split(b"\n")
            recs.append((_trim_cr(parts[0]), b"".join(_trim_cr(x) for x in parts[1:]), None))
        return FMT_FASTA, recs
    raise ValueError("invalid start")


def py_demux(text: bytes, barcodes: list[bytes], mismatch: int):
    """-> (assign list, per-bin canonical output bytes); candidate order = row order"""
    fmt, recs = py_records(text)
    k = len(barcodes[0])
    outs = [bytearray() for _ in range(len(barcodes) + 1)]
    asg = []
    for rid, s, q in recs:
        if len(s) < k:
            raise ValueError("short read")  # slice panic, src/demux.rs:54
        hit = next((i for i, bc in enumerate(barcodes) if py_bc_cmp(bc, s[:k], mismatch)), len(barcodes))
        asg.append(hit)
        if fmt == FMT_FASTQ:
            outs[hit] += b"@" + rid + b"\n" + s + b"\n+\n" + q + b"\n"
        else:
            outs[hit] += b">" + rid + b"\n" + s + b"\n"
    return asg, [bytes(o) for o in outs]
