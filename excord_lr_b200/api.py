"""ctypes binding of libexlr_cuda.so (include/exlr.h) — the host-side mirror of the reference
loop's interface: parameters are the reference Cli fields (reference src/main.rs:37-105), a
batch is the structure-of-arrays form of the records the loop iterates (src/main.rs:158),
results are the lines it writes (src/main.rs:395-766), in order.

There is no CPU fallback: if the CUDA library is missing or no B200 is visible, every entry
point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from .batch import EVENT_DTYPE, ExlrParams, HostBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libexlr_cuda.so")

EXLR_OPT_CIGAR_KERNEL = 1
EXLR_OPT_READS_PER_CTA = 2
EXLR_OPT_OVERLAP = 3
EXLR_OPT_K1_CTAS_PER_SM = 4
EXLR_OPT_K1_WAVES = 5
EXLR_OPT_STAGE_TIMING = 6
EXLR_OPT_TRACE = 7
EXLR_OPT_DEVICE_FORMAT = 8
EXLR_OPT_K1A_CTAS_PER_SM = 9
EXLR_OPT_LONG_RECORDS = 10
EXLR_OPT_VERBOSE_TEXT = 11
EXLR_OPT_K3_FOLD = 12
EXLR_OPT_GRAPH = 13
EXLR_OPT_WC_INPUT = 14
EXLR_OPT_K0_WALK = 15
EXLR_OPT_BGZF_CRC = 16
# see EXLR_OPT_CIGAR_KERNEL in include/exlr.h
CIGAR_KERNEL_AUTO, CIGAR_KERNEL_WARP, CIGAR_KERNEL_FLAT, CIGAR_KERNEL_SCREEN = 0, 1, 2, 3


class ExlrError(RuntimeError):
    def __init__(self, status: int, msg: str, err_read: int = -1):
        super().__init__(f"exlr status {status}: {msg}" + (f" (record {err_read})" if err_read >= 0 else ""))
        self.status, self.err_read = status, err_read


class ExlrCapacityError(ExlrError):
    """EXLR_ERR_CAPACITY from exlr_wait; .needed = the max_events that would have sufficed."""

    def __init__(self, needed: int):
        super().__init__(-4, f"event buffers too small, need max_events >= {needed}")
        self.needed = needed


class _Views(C.Structure):
    _fields_ = [("cigar", C.c_void_p), ("cigar_off", C.c_void_p), ("pos", C.c_void_p), ("tid", C.c_void_p),
                ("flag", C.c_void_p), ("mapq", C.c_void_p), ("sa_kind", C.c_void_p), ("sa_off", C.c_void_p),
                ("sa_bytes", C.c_void_p), ("max_reads", C.c_uint64), ("max_ops", C.c_uint64),
                ("max_sa_bytes", C.c_uint64), ("max_events", C.c_uint64)]


class _Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("err_read", C.c_uint32), ("n_reads", C.c_uint64), ("n_events", C.c_uint64),
                ("events", C.c_void_p), ("line_off", C.c_void_p), ("n_kept", C.c_uint64), ("n_sa_reads", C.c_uint64),
                ("n_cap_dropped", C.c_uint64), ("n_ops", C.c_uint64), ("n_err_lines", C.c_uint64)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("classify_ms", C.c_float), ("cigar_ms", C.c_float), ("sa_cigar_ms", C.c_float),
                ("sa_parse_ms", C.c_float), ("scan_ms", C.c_float), ("place_ms", C.c_float), ("kernels_ms", C.c_float),
                ("d2h_ms", C.c_float), ("launches", C.c_uint32), ("screen_ms", C.c_float),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("flagged_steps", "claimed_short", "claimed_warp", "claimed_long", "raw_events",
                                          "sa_events", "far_records", "text_bytes")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class BgzfBlock(C.Structure):
    _fields_ = [("comp_off", C.c_uint32), ("comp_len", C.c_uint32), ("ulen", C.c_uint32), ("crc32", C.c_uint32)]


class _BamViews(C.Structure):
    _fields_ = [("comp", C.c_void_p), ("blocks", C.c_void_p), ("max_comp_bytes", C.c_uint64), ("max_blocks", C.c_uint32),
                ("reserved", C.c_uint32), ("max_tail_bytes", C.c_uint64)]


class BamInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("bad_block", C.c_int32), ("n_blocks", C.c_uint32), ("reserved", C.c_uint32),
                ("n_reads", C.c_uint64), ("n_ops", C.c_uint64), ("n_sa_bytes", C.c_uint64), ("tail_off", C.c_uint64),
                ("u_bytes", C.c_uint64), ("comp_bytes", C.c_uint64), ("h2d_ms", C.c_float), ("inflate_ms", C.c_float),
                ("walk_ms", C.c_float), ("reserved2", C.c_float), ("t_ms", C.c_float * 4)]


_lib = None


def load_library() -> C.CDLL:
    """Loads libexlr_cuda.so; raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ExlrError(-2, f"{LIB_PATH} not built; run `make -C excord_lr_b200/csrc` (needs nvcc). "
                            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    lib.exlr_abi_version.restype = i32
    lib.exlr_device_count.restype = i32
    lib.exlr_params_default.argtypes = [C.POINTER(ExlrParams)]
    lib.exlr_create.argtypes = [C.POINTER(ExlrParams), i32, C.POINTER(C.c_char_p), i32, C.POINTER(vp)]
    lib.exlr_destroy.argtypes = [vp]
    lib.exlr_set_option.argtypes = [vp, i32, C.c_int64]
    lib.exlr_batch_alloc.argtypes = [vp, u64, u64, u64, u64, C.POINTER(vp)]
    lib.exlr_batch_free.argtypes = [vp]
    lib.exlr_batch_get_views.argtypes = [vp, C.POINTER(_Views)]
    lib.exlr_batch_grow.argtypes = [vp, u64]
    lib.exlr_submit.argtypes = [vp, u64]
    lib.exlr_upload.argtypes = [vp, u64]
    lib.exlr_submit_resident.argtypes = [vp]
    lib.exlr_wait.argtypes = [vp, C.POINTER(_Result)]
    lib.exlr_wait_resident.argtypes = [vp, C.POINTER(_Result)]
    lib.exlr_wait_text.argtypes = [vp, C.POINTER(_Result), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.exlr_get_timing.argtypes = [vp, C.POINTER(Timing)]
    lib.exlr_get_counters.argtypes = [vp, C.POINTER(Counters)]
    lib.exlr_get_counters.restype = i32
    lib.exlr_bam_batch_alloc.argtypes = [vp, u64, C.c_uint32, u64, u64, C.POINTER(vp)]
    lib.exlr_bam_get_views.argtypes = [vp, C.POINTER(_BamViews)]
    lib.exlr_bam_submit.argtypes = [vp, u64, C.c_uint32]
    lib.exlr_bam_walk.argtypes = [vp, vp, u64]
    lib.exlr_bam_extract.argtypes = [vp, C.POINTER(BamInfo)]
    lib.exlr_bam_download.argtypes = [vp, C.POINTER(_Views), vp, u64, vp]
    lib.exlr_bam_download_stream.argtypes = [vp, vp, u64, C.POINTER(u64)]
    lib.exlr_bam_download_stream.restype = i32
    for f in ("exlr_bam_batch_alloc", "exlr_bam_get_views", "exlr_bam_submit", "exlr_bam_walk", "exlr_bam_extract", "exlr_bam_download"):
        getattr(lib, f).restype = i32
    lib.exlr_format_lines.restype = C.c_int64
    lib.exlr_format_lines.argtypes = [vp, vp, C.POINTER(_Result), u64, u64, i32, C.c_char_p, vp, vp, u64]
    lib.exlr_strerror.restype = C.c_char_p
    lib.exlr_strerror.argtypes = [i32]
    lib.exlr_last_cuda_error.restype = C.c_char_p
    for f in ("exlr_create", "exlr_set_option", "exlr_batch_alloc", "exlr_batch_get_views", "exlr_batch_grow", "exlr_submit", "exlr_upload",
              "exlr_submit_resident", "exlr_wait", "exlr_wait_resident", "exlr_wait_text", "exlr_get_timing"):
        getattr(lib, f).restype = i32
    _lib = lib
    return lib


def _check(rc: int, err_read: int = -1):
    if rc != 0:
        lib = load_library()
        msg = lib.exlr_strerror(rc).decode()
        if rc == -2:
            msg += ": " + lib.exlr_last_cuda_error().decode()
        raise ExlrError(rc, msg, err_read)


class Result:
    """Outcome of one batch: ordered events (EVENT_DTYPE), per-record line offsets, counters."""

    def __init__(self, raw: _Result, copy: bool, resident: bool = False):
        self.status, self.err_read = int(raw.status), int(raw.err_read)
        self.n_reads, self.n_events = int(raw.n_reads), int(raw.n_events)
        self.n_kept, self.n_sa_reads = int(raw.n_kept), int(raw.n_sa_reads)
        self.n_cap_dropped, self.n_ops = int(raw.n_cap_dropped), int(raw.n_ops)
        self.n_err_lines = int(raw.n_err_lines)
        self._raw = raw
        if resident:
            self.events, self.line_off = None, None
            return
        ev = np.frombuffer((C.c_char * (self.n_events * 48)).from_address(raw.events), EVENT_DTYPE) if self.n_events \
            else np.zeros(0, EVENT_DTYPE)
        lo = np.frombuffer((C.c_char * ((self.n_reads + 1) * 4)).from_address(raw.line_off), np.uint32)
        self.events = ev.copy() if copy else ev
        self.line_off = lo.copy() if copy else lo


    def n_valid_lines(self) -> int:
        """Lines the reference writes for this batch: all of them, or on a per-record panic (status <= -10) the lines of
        the records before err_read plus the ones that record itself had written by then."""
        if self.status > -10:
            return self.n_events
        return int(self.line_off[self.err_read]) + self.n_err_lines


class DeviceBatch:
    """exlr_batch: pinned SoA views + device buffers + a stream."""

    def __init__(self, ex: "Extractor", max_reads: int, max_ops: int, max_sa_bytes: int, max_events: int = 0):
        self.ex, self.lib = ex, ex.lib
        h = C.c_void_p()
        _check(self.lib.exlr_batch_alloc(ex.handle, max_reads, max_ops, max_sa_bytes, max_events, C.byref(h)))
        self.handle = h
        v = _Views()
        _check(self.lib.exlr_batch_get_views(h, C.byref(v)))
        self.max_reads, self.max_ops, self.max_sa_bytes, self.max_events = (int(v.max_reads), int(v.max_ops),
                                                                            int(v.max_sa_bytes), int(v.max_events))

        def view(ptr, n, dt):
            return np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(ptr), dt)
        self.cigar = view(v.cigar, max_ops + 4, np.uint32)
        self.cigar_off = view(v.cigar_off, max_reads + 1, np.uint64)
        self.pos = view(v.pos, max_reads, np.int32)
        self.tid = view(v.tid, max_reads, np.int32)
        self.flag = view(v.flag, max_reads, np.uint16)
        self.mapq = view(v.mapq, max_reads, np.uint8)
        self.sa_kind = view(v.sa_kind, max_reads, np.uint8)
        self.sa_off = view(v.sa_off, max_reads + 1, np.uint32)
        self.sa_bytes = view(v.sa_bytes, max_sa_bytes + 16, np.uint8)
        self.n_reads = 0
        self._qn = None

    def fill(self, hb: HostBatch):
        """The packer step: copy a HostBatch into the pinned views."""
        n, c, a = hb.n_reads, hb.n_ops, hb.n_sa_bytes
        if n > self.max_reads or c > self.max_ops or a > self.max_sa_bytes:
            raise ExlrError(-4, "batch exceeds its allocated capacity")
        self.cigar[:c] = hb.cigar[:c]
        self.cigar_off[:n + 1] = hb.cigar_off
        self.pos[:n] = hb.pos
        self.tid[:n] = hb.tid
        self.flag[:n] = hb.flag
        self.mapq[:n] = hb.mapq
        self.sa_kind[:n] = hb.sa_kind
        self.sa_off[:n + 1] = hb.sa_off
        self.sa_bytes[:a] = hb.sa_bytes[:a]
        self.n_reads = n
        self._qn = None
        return self

    def grow(self, max_events: int):
        """Larger event buffers, same packed records: the answer to ExlrCapacityError (then submit again)."""
        _check(self.lib.exlr_batch_grow(self.handle, max_events))
        self.max_events = max(self.max_events, max_events)

    def submit(self, n_reads: Optional[int] = None):
        _check(self.lib.exlr_submit(self.handle, self.n_reads if n_reads is None else n_reads))

    def upload(self, n_reads: Optional[int] = None):
        _check(self.lib.exlr_upload(self.handle, self.n_reads if n_reads is None else n_reads))

    def submit_resident(self):
        _check(self.lib.exlr_submit_resident(self.handle))

    def wait(self, copy: bool = True, raise_on_record_error: bool = False) -> Result:
        raw = _Result()
        rc = self.lib.exlr_wait(self.handle, C.byref(raw))
        if rc == -4:                                    # event buffers too small: n_events = capacity needed
            raise ExlrCapacityError(int(raw.n_events))
        if rc != 0 and (rc > -10 or raise_on_record_error):
            _check(rc, int(raw.err_read) if rc <= -10 else -1)
        return Result(raw, copy)

    def wait_text(self):
        """With EXLR_OPT_DEVICE_FORMAT: (Result without events, the formatted non-verbose lines as bytes)."""
        raw, ptr, n = _Result(), C.c_void_p(), C.c_uint64()
        rc = self.lib.exlr_wait_text(self.handle, C.byref(raw), C.byref(ptr), C.byref(n))
        if rc == -4:
            raise ExlrCapacityError(int(raw.n_events))
        if rc != 0 and rc > -10:
            _check(rc)
        return Result(raw, False, resident=True), (C.string_at(ptr.value, n.value) if n.value else b"")

    def wait_resident(self) -> Result:
        raw = _Result()
        rc = self.lib.exlr_wait_resident(self.handle, C.byref(raw))
        if rc != 0 and rc > -10:
            _check(rc)
        return Result(raw, False, resident=True)

    def timing(self) -> Timing:
        t = Timing()
        _check(self.lib.exlr_get_timing(self.handle, C.byref(t)))
        return t

    def counters(self) -> Counters:
        c = Counters()
        _check(self.lib.exlr_get_counters(self.handle, C.byref(c)))
        return c

    def format_lines(self, res: Result, verbose: bool = False, qnames: Optional[Sequence[str]] = None,
                     ev_begin: int = 0, ev_end: Optional[int] = None) -> bytes:
        """The bytes the reference writes for these events (reference src/utils.rs:196-283)."""
        ev_end = res.n_events if ev_end is None else ev_end
        qb, qo = None, None
        if verbose:
            names = list(qnames) if qnames is not None else ["r%09d" % i for i in range(res.n_reads)]
            enc = [s.encode() for s in names]
            qo = np.zeros(len(enc) + 1, np.uint32)
            qo[1:] = np.cumsum([len(e) for e in enc])
            qb = b"".join(enc)
        qo_p = qo.ctypes.data if qo is not None else None
        need = self.lib.exlr_format_lines(self.ex.handle, self.handle, C.byref(res._raw), ev_begin, ev_end, int(verbose),
                                          qb, qo_p, None, 0)
        if need < 0:
            _check(int(need))
        buf = C.create_string_buffer(int(need) + 1)
        self.lib.exlr_format_lines(self.ex.handle, self.handle, C.byref(res._raw), ev_begin, ev_end, int(verbose),
                                   qb, qo_p, C.cast(buf, C.c_void_p), need)
        return buf.raw[:need]

    def free(self):
        if self.handle:
            self.lib.exlr_batch_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def bgzf_blocks(data: bytes):
    """The host's whole share of reading a BGZF file: hop over the block headers.  -> [(comp_off, comp_len, ulen)] of the
    complete blocks in `data`, and the number of bytes they span (SAMv1 4.1: gzip header with the BC extra subfield)."""
    out, at, n = [], 0, len(data)
    while at + 18 <= n:
        if data[at] != 31 or data[at + 1] != 139 or data[at + 2] != 8 or not (data[at + 3] & 4):
            raise ExlrError(-7, "not a BGZF block")
        xlen = int.from_bytes(data[at + 10:at + 12], "little")
        if at + 12 + xlen > n:
            break
        bsize, i = -1, at + 12
        while i + 4 <= at + 12 + xlen:
            slen = int.from_bytes(data[i + 2:i + 4], "little")
            if data[i] == 66 and data[i + 1] == 67 and slen == 2:
                bsize = int.from_bytes(data[i + 4:i + 6], "little")
            i += 4 + slen
        if bsize < 0:
            raise ExlrError(-7, "BGZF block without BC field")
        total = bsize + 1
        if at + total > n:
            break
        out.append((at + 12 + xlen, total - xlen - 20, int.from_bytes(data[at + total - 4:at + total], "little")))
        at += total
    return out, at


class BamBatch(DeviceBatch):
    """A batch whose records come from BGZF-compressed BAM bytes decoded on the device (exlr_bam_*)."""

    def __init__(self, ex: "Extractor", max_comp_bytes: int, max_blocks: int, max_events: int = 0, max_tail_bytes: int = 1 << 20):
        self.ex, self.lib = ex, ex.lib
        h = C.c_void_p()
        _check(self.lib.exlr_bam_batch_alloc(ex.handle, max_comp_bytes, max_blocks, max_tail_bytes, max_events, C.byref(h)))
        self.handle = h
        v = _BamViews()
        _check(self.lib.exlr_bam_get_views(h, C.byref(v)))
        self.max_comp, self.max_blocks = int(v.max_comp_bytes), int(v.max_blocks)
        self.comp = np.frombuffer((C.c_char * self.max_comp).from_address(v.comp), np.uint8)
        self.blocks = (BgzfBlock * self.max_blocks).from_address(v.blocks)
        self.n_reads = 0

    def load(self, data: bytes, blocks):
        """Chunk bytes + their block table (bgzf_blocks) into the pinned views, then H2D + inflate (asynchronous)."""
        self.comp[:len(data)] = np.frombuffer(data, np.uint8)
        for i, (co, cl, ul) in enumerate(blocks):
            if co + cl + 4 > len(data):
                raise ExlrError(-1, "the chunk must hold whole BGZF blocks, footers included (CRC32)")
            self.blocks[i] = BgzfBlock(co, cl, ul, int.from_bytes(data[co + cl:co + cl + 4], "little"))
        _check(self.lib.exlr_bam_submit(self.handle, len(data), len(blocks)))

    def walk(self, start_off: int = 0, prev: "BamBatch" = None):
        """Record walk + gather (asynchronous).  prev: the batch of the previous chunk (already extracted): the bytes its walk left
        over -- its partial last record -- are copied in front of this chunk's on the device; start_off counts from there."""
        _check(self.lib.exlr_bam_walk(self.handle, prev.handle if prev is not None else None, start_off))

    def extract(self) -> BamInfo:
        info = BamInfo()
        rc = self.lib.exlr_bam_extract(self.handle, C.byref(info))
        if rc not in (0, -7, -8):
            _check(rc)
        self.n_reads = int(info.n_reads)
        return info

    def inflated(self) -> bytes:
        """The chunk's inflated byte stream (raises ExlrError -7 if a block did not inflate)."""
        n = C.c_uint64()
        _check(self.lib.exlr_bam_download_stream(self.handle, None, 0, C.byref(n)))
        buf = np.zeros(max(1, n.value), np.uint8)
        _check(self.lib.exlr_bam_download_stream(self.handle, buf.ctypes.data, buf.size, C.byref(n)))
        return buf[:n.value].tobytes()

    def download(self, ref_names, max_reads, max_ops, max_sa) -> HostBatch:
        """The decoded records as a HostBatch (for comparison with what the host reader packs)."""
        n, c, a = max(1, max_reads), max(1, max_ops), max(1, max_sa)
        arr = dict(cigar=np.zeros(c, np.uint32), cigar_off=np.zeros(n + 1, np.uint64), pos=np.zeros(n, np.int32),
                   tid=np.zeros(n, np.int32), flag=np.zeros(n, np.uint16), mapq=np.zeros(n, np.uint8), sa_kind=np.zeros(n, np.uint8),
                   sa_off=np.zeros(n + 1, np.uint32), sa_bytes=np.zeros(a, np.uint8))
        v = _Views(*[arr[k].ctypes.data for k in ("cigar", "cigar_off", "pos", "tid", "flag", "mapq", "sa_kind", "sa_off", "sa_bytes")],
                   n, c, a, 0)
        qn = np.zeros(256 * n, np.uint8)
        qo = np.zeros(n + 1, np.uint32)
        _check(self.lib.exlr_bam_download(self.handle, C.byref(v), qn.ctypes.data, qn.size, qo.ctypes.data))
        k = max_reads
        names = [qn[int(qo[i]):int(qo[i + 1])].tobytes().decode("latin-1") for i in range(k)]
        return HostBatch(arr["cigar"][:int(arr["cigar_off"][k])], arr["cigar_off"][:k + 1], arr["pos"][:k], arr["tid"][:k],
                         arr["flag"][:k], arr["mapq"][:k], arr["sa_kind"][:k], arr["sa_off"][:k + 1],
                         arr["sa_bytes"][:int(arr["sa_off"][k])], list(ref_names), names)


class Extractor:
    """exlr_ctx: one per GPU."""

    def __init__(self, params: ExlrParams, ref_names: Sequence[str], device: int = 0):
        self.lib = load_library()
        self.params, self.ref_names, self.device = params, list(ref_names), device
        arr = (C.c_char_p * max(1, len(self.ref_names)))()
        for i, n in enumerate(self.ref_names):
            arr[i] = n.encode("latin-1")
        h = C.c_void_p()
        _check(self.lib.exlr_create(C.byref(params), device, arr, len(self.ref_names), C.byref(h)))
        self.handle = h

    def set_option(self, option: int, value: int):
        _check(self.lib.exlr_set_option(self.handle, option, value))

    def alloc_batch(self, max_reads: int, max_ops: int, max_sa_bytes: int, max_events: int = 0) -> DeviceBatch:
        return DeviceBatch(self, max_reads, max_ops, max_sa_bytes, max_events)

    def batch_for(self, hb: HostBatch, max_events: int = 0) -> DeviceBatch:
        return self.alloc_batch(max(1, hb.n_reads), hb.n_ops, hb.n_sa_bytes, max_events).fill(hb)

    def close(self):
        if self.handle:
            self.lib.exlr_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def extract(hb: HostBatch, params: ExlrParams, device: int = 0, cigar_kernel: int = CIGAR_KERNEL_AUTO,
            reads_per_cta: int = 0, verbose: bool = False, max_events: int = 0, grow: bool = True, device_format: bool = False, long_records: int = 0):
    """One-shot: host batch -> (Result, formatted lines).  Host buffers in, host buffers out.
    If the event buffers turn out too small the batch is re-run once with the size the device reported.
    device_format: also format the (non-verbose) lines on the device; they come back as Result.device_text."""
    ex = Extractor(params, hb.ref_names, device)
    try:
        ex.set_option(EXLR_OPT_CIGAR_KERNEL, cigar_kernel)
        ex.set_option(EXLR_OPT_READS_PER_CTA, reads_per_cta)
        ex.set_option(EXLR_OPT_DEVICE_FORMAT, int(device_format))
        ex.set_option(EXLR_OPT_LONG_RECORDS, long_records)
        b = ex.batch_for(hb, max_events)
        try:
            for attempt in range(5):
                b.submit()
                try:
                    dtext = b.wait_text()[1] if device_format else None
                    res = b.wait()
                    res.device_text = dtext
                except ExlrCapacityError as e:
                    if not grow or attempt == 4:
                        raise
                    b.grow(e.needed + 16)                   # same packed records, larger event buffers
                    continue
                text = b.format_lines(res, verbose, hb.qnames, 0, res.n_valid_lines())
                return res, text
        finally:
            b.free()
    finally:
        ex.close()
