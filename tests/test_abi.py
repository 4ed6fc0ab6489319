"""No-GPU checks of the boundary: the library loads, exports every symbol the header declares, struct sizes
match, and the compute entry points fail loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re

import pytest

from excord_lr_b200 import api
from excord_lr_b200.batch import EVENT_DTYPE, ExlrParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "exlr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(exlr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = api.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/exlr.h but not exported"
    assert lib.exlr_abi_version() == 2


def test_struct_layouts():
    assert C.sizeof(ExlrParams) == 40 and EVENT_DTYPE.itemsize == 48
    assert C.sizeof(api._Result) == 80 and C.sizeof(api.Timing) == 64 and C.sizeof(api._Views) == 104
    assert C.sizeof(api.BamInfo) == 96 and C.sizeof(api.BgzfBlock) == 16 and C.sizeof(api._BamViews) == 40 and C.sizeof(api.Counters) == 64
    lib = api.load_library()
    p = ExlrParams()
    lib.exlr_params_default(C.byref(p))
    assert (p.mapq, p.exclude_flag, p.indel_min, p.merge_min, p.ins_clip_min, p.max_pct_overlap, p.max_supp_alignm) == \
        (1, 1796, 50, 5, 1000, 0.0, 4)          # reference src/main.rs:47-96


def test_no_cpu_fallback_without_device():
    lib = api.load_library()
    if lib.exlr_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.ExlrError) as e:
        api.Extractor(ExlrParams.make(), ["chr1"])
    assert e.value.status == -2


def test_product_does_not_touch_the_oracle():
    # the shipped package must never import, load or link anything under oracle/
    pkg = os.path.join(ROOT, "excord_lr_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|oracle_c|libexlr_oracle|exlr_oracle|oracle/")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not pat.search(txt), f"{os.path.join(dp, f)} references the oracle"
