"""Error behaviour of the C ABI (include/sabreur_gpu.h) on a GPU box: bad arguments come back as negative codes with a
message in sbr_last_error(), never as a crash, and a context stays usable afterwards.  The reference has no FFI seam;
these are the contract of the boundary that replaces `se_demux` / `pe_demux` (src/demux.rs:13-20, 71-79), whose own
failures are `anyhow` errors (src/demux.rs:22, 50)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

E_INVALID_ARG, E_UNSUPPORTED, E_BUSY = -1, -4, -5


@pytest.fixture(scope="module")
def lib():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test without a CUDA device (there is no CPU fallback to test)")
    from sabreur_b200 import _ffi
    return _ffi, _ffi.load()


def _cfg(ffi, rows=(b"ACGTAC", b"TTGACA"), lens=None, **kw):
    tab = np.frombuffer(b"".join(r.ljust(len(rows[0]), b"A") for r in rows), dtype=np.uint8).copy()
    cfg = ffi.SbrConfig(device=0, fmt=ffi.FMT_AUTO, n_barcodes=len(rows), bc_len=len(rows[0]), barcodes=tab.ctypes.data,
                        bc_lens=None, mismatch=1, n_slots=2, batch_bytes=1 << 20, max_records=0, flags=0)
    keep = [tab]
    if lens is not None:
        la = np.asarray(lens, dtype=np.uint32)
        cfg.bc_lens = la.ctypes.data
        keep.append(la)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg, keep


def _create(ffi, L, **kw):
    cfg, keep = _cfg(ffi, **kw)
    h = C.c_void_p()
    rc = L.sbr_create(C.byref(h), C.byref(cfg))
    return rc, h


@pytest.mark.parametrize("kw,code", [
    (dict(n_barcodes=0), E_INVALID_ARG),
    (dict(bc_len=0), E_INVALID_ARG),
    (dict(barcodes=None), E_INVALID_ARG),
    (dict(device=99), E_INVALID_ARG),
    (dict(device=-1), E_INVALID_ARG),
    (dict(batch_bytes=0), E_INVALID_ARG),
    (dict(batch_bytes=(1 << 30) + 1), E_INVALID_ARG),
    (dict(n_slots=0), E_INVALID_ARG),
    (dict(n_slots=65), E_INVALID_ARG),
    (dict(fmt=3), E_INVALID_ARG),
    (dict(mismatch=256), E_INVALID_ARG),       # a u8 in the reference (src/cli.rs:33)
    (dict(lens=[6, 7]), E_INVALID_ARG),        # a row longer than the window
    (dict(lens=[5, 6]), E_INVALID_ARG),        # row 0 defines bc_len (src/demux.rs:31-34)
    (dict(bc_len=65), E_UNSUPPORTED),
    (dict(n_barcodes=6001), E_UNSUPPORTED),
])
def test_create_rejects_bad_configurations(lib, kw, code):
    ffi, L = lib
    rc, h = _create(ffi, L, **kw)
    assert rc == code and not h.value
    assert L.sbr_last_error(None)  # the message of a failed create is kept per thread


def test_null_arguments(lib):
    ffi, L = lib
    cfg, _keep = _cfg(ffi)
    assert L.sbr_create(None, C.byref(cfg)) == E_INVALID_ARG
    h = C.c_void_p()
    assert L.sbr_create(C.byref(h), None) == E_INVALID_ARG
    r = ffi.SbrResult()
    assert L.sbr_wait(None, 0, C.byref(r)) == E_INVALID_ARG
    assert L.sbr_submit(None, 0, 0, 0, 0, 1) == E_INVALID_ARG
    L.sbr_destroy(None)  # a no-op


def test_slot_protocol_violations_leave_the_context_usable(lib):
    ffi, L = lib
    rc, h = _create(ffi, L)
    assert rc == 0
    try:
        r = ffi.SbrResult()
        ptr, cap = C.c_void_p(), C.c_uint64()
        assert L.sbr_slot_buffer(h, 2, C.byref(ptr), C.byref(cap)) == E_INVALID_ARG      # only slots 0 and 1 exist
        assert L.sbr_wait(h, 0, C.byref(r)) == E_INVALID_ARG                              # nothing in flight
        assert b"nothing in flight" in L.sbr_last_error(h)
        assert L.sbr_slot_buffer(h, 0, C.byref(ptr), C.byref(cap)) == 0 and cap.value == 1 << 20
        text = b"@a\nACGTACGG\n+\nIIIIIIII\n@b\nTTGACAGG\n+\nIIIIIIII\n@c\nGGGGGGGG\n+\nIIIIIIII\n"
        C.memmove(ptr.value, text, len(text))
        assert L.sbr_submit(h, 0, (1 << 20) - 10, 11, 0, 1) == E_INVALID_ARG              # runs past the slot
        assert L.sbr_submit(h, 7, 0, len(text), 0, 1) == E_INVALID_ARG
        assert L.sbr_submit(h, 0, 0, len(text), 5, 1) == 0
        assert L.sbr_submit(h, 0, 0, len(text), 6, 1) == E_BUSY                            # not waited for yet
        assert L.sbr_wait(h, 0, None) == E_INVALID_ARG
        assert L.sbr_wait(h, 0, C.byref(r)) == 0
        assert (r.batch_seq, r.n_records, r.status_flags & ffi.ST_ERROR_MASK, r.consumed) == (5, 3, 0, len(text))
        assign = np.ctypeslib.as_array(C.cast(r.assign, C.POINTER(C.c_uint16)), shape=(3,))
        assert assign.tolist() == [0, 1, 2]
        assert L.sbr_wait(h, 0, C.byref(r)) == E_INVALID_ARG                              # a batch is waited for once
        # a stream that starts with neither '>' nor '@' is refused like needletail does; an empty final batch is fine
        C.memmove(ptr.value, b"ACGT\n", 5)
        assert L.sbr_submit(h, 0, 0, 5, 0, 1) == 0 and L.sbr_wait(h, 0, C.byref(r)) == 0
        assert r.n_records == 0 and r.status_flags & ffi.ST_INVALID_START
        assert L.sbr_submit(h, 1, 0, 0, 1, 1) == 0 and L.sbr_wait(h, 1, C.byref(r)) == 0
        assert r.n_records == 0 and r.status_flags == 0 and r.consumed == 0
        # and the context still works
        C.memmove(ptr.value, text, len(text))
        assert L.sbr_submit(h, 0, 0, len(text), 9, 1) == 0 and L.sbr_wait(h, 0, C.byref(r)) == 0
        assert r.n_records == 3 and r.batch_seq == 9
    finally:
        L.sbr_destroy(h)


def test_device_resident_entry_point_checks_its_arguments(lib):
    ffi, L = lib
    rc, h = _create(ffi, L, fmt=ffi.FMT_FASTQ)
    assert rc == 0
    d = C.c_void_p()
    assert L.sbr_dev_alloc(0, 4096, C.byref(d)) == 0
    try:
        assert L.sbr_demux_device(h, 0, None, 100, 1, None) == E_INVALID_ARG
        assert L.sbr_demux_device(h, 0, d, 0, 1, None) == E_INVALID_ARG
        assert L.sbr_demux_device(h, 0, d.value + 1, 100, 1, None) == E_INVALID_ARG       # 16-byte alignment
        assert L.sbr_demux_device(h, 0, d, (1 << 30) + 1, 1, None) == E_INVALID_ARG
        assert L.sbr_demux_device(h, 0, d, 8 << 20, 1, None) == E_INVALID_ARG             # more than batch_bytes
        assert L.sbr_demux_device(h, 9, d, 100, 1, None) == E_INVALID_ARG
    finally:
        L.sbr_dev_free(0, d)
        L.sbr_destroy(h)
    rc, h = _create(ffi, L)  # fmt AUTO: the format of a device-resident batch cannot be sniffed on the host
    assert rc == 0
    try:
        assert L.sbr_dev_alloc(0, 4096, C.byref(d)) == 0
        assert L.sbr_demux_device(h, 0, d, 100, 1, None) == E_INVALID_ARG
        L.sbr_dev_free(0, d)
    finally:
        L.sbr_destroy(h)
