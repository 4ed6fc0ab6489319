"""excord_lr_b200 — B200-native (sm_100a) signal extraction for long-read BAMs, a drop-in for the
per-record loop of excord-lr (reference src/main.rs:158-770).

The compute path is the CUDA library csrc/libexlr_cuda.so behind the C ABI of include/exlr.h;
this package is the thin host mirror (ctypes).  No CPU fallback exists.
"""
from .batch import EVENT_DTYPE, ExlrParams, HostBatch, pack_records  # noqa: F401
from .api import (CIGAR_KERNEL_AUTO, CIGAR_KERNEL_FLAT, CIGAR_KERNEL_SCREEN, CIGAR_KERNEL_WARP, DeviceBatch, ExlrCapacityError, ExlrError, Extractor, Result, extract,  # noqa: F401
                  load_library)

__all__ = ["EVENT_DTYPE", "ExlrParams", "HostBatch", "pack_records", "Extractor", "DeviceBatch", "Result", "ExlrError", "ExlrCapacityError",
           "extract", "load_library", "CIGAR_KERNEL_AUTO", "CIGAR_KERNEL_FLAT", "CIGAR_KERNEL_SCREEN", "CIGAR_KERNEL_WARP"]
