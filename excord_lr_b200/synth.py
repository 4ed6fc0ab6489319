"""ctypes front end of the deterministic synthetic-batch generator (synth/exlr_synth.cpp).

The BASELINE.json configs are exposed as `config(i, scale)`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .batch import HostBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "synth", "exlr_synth.cpp")
_SO = os.path.join(_HERE, "synth", "libexlr_synth.so")


class _Out(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_ops", C.c_uint64), ("n_sa_bytes", C.c_uint64),
                ("cigar", C.POINTER(C.c_uint32)), ("cigar_off", C.POINTER(C.c_uint64)), ("pos", C.POINTER(C.c_int32)),
                ("tid", C.POINTER(C.c_int32)), ("flag", C.POINTER(C.c_uint16)), ("mapq", C.POINTER(C.c_uint8)),
                ("sa_kind", C.POINTER(C.c_uint8)), ("sa_off", C.POINTER(C.c_uint32)), ("sa_bytes", C.POINTER(C.c_uint8)),
                ("qid", C.POINTER(C.c_uint64))]


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _SO, _SRC])
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.exlr_synth_generate.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.POINTER(_Out)]
        _lib.exlr_synth_ref_name.restype = C.c_char_p
        _lib.exlr_synth_ref_len.restype = C.c_int64
    return _lib


def ref_names():
    lib = _load()
    return [lib.exlr_synth_ref_name(i).decode() for i in range(lib.exlr_synth_n_ref())]


def ref_lens():
    lib = _load()
    return [int(lib.exlr_synth_ref_len(i)) for i in range(lib.exlr_synth_n_ref())]


PROFILE_HIFI, PROFILE_ONT, PROFILE_SPLIT = 0, 1, 2


def generate(profile: int, seed: int, n_molecules: int, chr20_only: bool = False, n_ultra: int = 0) -> HostBatch:
    lib = _load()
    o = _Out()
    rc = lib.exlr_synth_generate(profile, seed, n_molecules, int(chr20_only), n_ultra, C.byref(o))
    if rc != 0:
        raise RuntimeError(f"exlr_synth_generate failed: {rc}")
    try:
        n, c, a = o.n_reads, o.n_ops, o.n_sa_bytes

        def arr(p, cnt, dt):
            if cnt == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(p, shape=(cnt,)).astype(dt, copy=True)
        qid = arr(o.qid, n, np.uint64)
        hb = HostBatch(arr(o.cigar, c, np.uint32), arr(o.cigar_off, n + 1, np.uint64), arr(o.pos, n, np.int32),
                       arr(o.tid, n, np.int32), arr(o.flag, n, np.uint16), arr(o.mapq, n, np.uint8),
                       arr(o.sa_kind, n, np.uint8), arr(o.sa_off, n + 1, np.uint32), arr(o.sa_bytes, a, np.uint8),
                       ref_names(), None)
        hb.qid = qid
        return hb
    finally:
        lib.exlr_synth_free(C.byref(o))


def with_qnames(hb: HostBatch) -> HostBatch:
    hb.qnames = ["r%09d" % int(q) for q in hb.qid]
    return hb


# BASELINE.json configs[i] -> (profile, seed, n_molecules, chr20_only, n_ultra, params kwargs)
CONFIGS = {
    0: dict(name="c1_hifi_10k_chr20", profile=PROFILE_HIFI, seed=1001, n=10_000, chr20=True, ultra=0,
            params=dict(mapq=1, exclude_flag=1796, indel_min=50, merge_min=5, max_pct_overlap=0.8, max_supp_alignm=4)),
    1: dict(name="c2_hifi_1M", profile=PROFILE_HIFI, seed=1002, n=1_000_000, chr20=False, ultra=0,
            params=dict(mapq=1, exclude_flag=1796, indel_min=50, merge_min=5, max_pct_overlap=0.8, max_supp_alignm=4)),
    2: dict(name="c3_ont_500k", profile=PROFILE_ONT, seed=1003, n=500_000, chr20=False, ultra=100,
            params=dict(indel_min=30)),
    3: dict(name="c4_split_2M", profile=PROFILE_SPLIT, seed=1004, n=670_000, chr20=False, ultra=0,
            params=dict(split_only=True, max_supp_alignm=4)),
    4: dict(name="c5_hifi_6M", profile=PROFILE_HIFI, seed=1005, n=6_000_000, chr20=False, ultra=0,
            params=dict(mapq=1, exclude_flag=1796, indel_min=50, merge_min=5, max_pct_overlap=0.8, max_supp_alignm=4)),
}


def config(i: int, scale: float = 1.0) -> HostBatch:
    """Synthetic batch of BASELINE.json configs[i]; scale < 1 shrinks the molecule count."""
    c = CONFIGS[i]
    n = max(1, int(c["n"] * scale))
    ultra = c["ultra"] if scale >= 1.0 else min(c["ultra"], max(0, int(c["ultra"] * scale)))
    return generate(c["profile"], c["seed"], n, c["chr20"], ultra)
