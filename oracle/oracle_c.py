"""ctypes loader of the C oracle (oracle/exlr_oracle.c) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  The product package (excord_lr_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from excord_lr_b200.batch import EVENT_DTYPE, ExlrParams, HostBatch  # noqa: E402  (shared data layout only)

_SO = os.path.join(_HERE, "libexlr_oracle.so")
_SRC = os.path.join(_HERE, "exlr_oracle.c")


class _Out(C.Structure):
    _fields_ = [("events", C.c_void_p), ("n_events", C.c_uint64), ("cap_events", C.c_uint64),
                ("line_off", C.POINTER(C.c_uint32)), ("status", C.c_int32), ("err_read", C.c_uint32),
                ("n_kept", C.c_uint64), ("n_sa_reads", C.c_uint64), ("n_cap_dropped", C.c_uint64), ("n_ops", C.c_uint64),
                ("n_err_lines", C.c_uint64)]


def build(force: bool = False) -> str:
    hdr = os.path.join(_HERE, "..", "include", "exlr.h")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(_SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", _SO, _SRC])
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.exlr_oracle_run.restype = C.c_int
        _lib.exlr_oracle_run.argtypes = [C.POINTER(ExlrParams), C.POINTER(C.c_char_p), C.c_int, C.c_uint64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64,
                                         C.POINTER(_Out)]
        _lib.exlr_oracle_format.restype = C.c_int64
        _lib.exlr_oracle_format.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_char_p, C.c_void_p, C.c_char_p, C.c_uint64]
        _lib.exlr_oracle_free.argtypes = [C.POINTER(_Out)]
    return _lib


def _names(ref_names):
    arr = (C.c_char_p * max(1, len(ref_names)))()
    for i, n in enumerate(ref_names):
        arr[i] = n.encode()
    return arr


class OracleResult:
    """events: every line the reference writes, i.e. for a failing run the lines of the records before err_read plus the
    n_err_lines lines the failing record itself had written before it panicked."""

    def __init__(self, status, err_read, events, line_off, n_kept, n_sa_reads, n_cap_dropped, n_ops, n_err_lines=0):
        self.status, self.err_read, self.events, self.line_off = status, err_read, events, line_off
        self.n_kept, self.n_sa_reads, self.n_cap_dropped, self.n_ops = n_kept, n_sa_reads, n_cap_dropped, n_ops
        self.n_err_lines = n_err_lines


def run(hb: HostBatch, params: ExlrParams, merge_mode: int = 1, r_begin: int = 0, r_end: int | None = None) -> OracleResult:
    lib = _load()
    n = hb.n_reads
    r_end = n if r_end is None else r_end
    o = _Out()
    names = _names(hb.ref_names)
    sab = hb.sa_bytes if hb.sa_bytes.size else np.zeros(1, np.uint8)
    cig = hb.cigar if hb.cigar.size else np.zeros(1, np.uint32)
    lib.exlr_oracle_run(C.byref(params), names, len(hb.ref_names), n, cig.ctypes.data, hb.cigar_off.ctypes.data,
                        hb.pos.ctypes.data, hb.tid.ctypes.data, hb.flag.ctypes.data, hb.mapq.ctypes.data,
                        hb.sa_kind.ctypes.data, hb.sa_off.ctypes.data, sab.ctypes.data, merge_mode, r_begin, r_end,
                        C.byref(o))
    try:
        ne = int(o.n_events)
        ev = np.zeros(ne, EVENT_DTYPE)
        if ne:
            C.memmove(ev.ctypes.data, o.events, ne * EVENT_DTYPE.itemsize)
        lo = np.ctypeslib.as_array(o.line_off, shape=(r_end - r_begin + 1,)).copy()
        return OracleResult(int(o.status), int(o.err_read), ev, lo, int(o.n_kept), int(o.n_sa_reads),
                            int(o.n_cap_dropped), int(o.n_ops), int(o.n_err_lines))
    finally:
        lib.exlr_oracle_free(C.byref(o))


def run_sharded(hb: HostBatch, params: ExlrParams, threads: int) -> list:
    """The loop body over `threads` contiguous record shards at once (ctypes drops the GIL).
    The reference's loop is single-threaded (src/main.rs:158); this exists only to give the
    CPU baseline every host core.  Returns the per-shard results in order."""
    n = hb.n_reads
    # balance by CIGAR ops
    targets = np.linspace(0, float(hb.cigar_off[-1]), threads + 1)
    cuts = np.searchsorted(hb.cigar_off[:-1].astype(np.float64), targets[1:-1]).tolist()
    bounds = [0] + cuts + [n]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        futs = [ex.submit(run, hb, params, 0, bounds[i], bounds[i + 1]) for i in range(threads)]
        return [f.result() for f in futs]


def format_lines(hb: HostBatch, events: np.ndarray, verbose: bool = False) -> bytes:
    lib = _load()
    names = _names(hb.ref_names)
    sab = hb.sa_bytes if hb.sa_bytes.size else np.zeros(1, np.uint8)
    qn, qo = (hb.qname_blob() if verbose else (b"", np.zeros(1, np.uint32)))
    ev = np.ascontiguousarray(events)
    need = lib.exlr_oracle_format(ev.ctypes.data, len(ev), names, sab.ctypes.data, hb.flag.ctypes.data, int(verbose),
                                  qn, qo.ctypes.data, None, 0)
    buf = C.create_string_buffer(int(need) + 1)
    lib.exlr_oracle_format(ev.ctypes.data, len(ev), names, sab.ctypes.data, hb.flag.ctypes.data, int(verbose),
                           qn, qo.ctypes.data, buf, need)
    return buf.raw[:need]
