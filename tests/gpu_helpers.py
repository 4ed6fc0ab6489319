"""GPU parity checker: the CUDA path (through the C ABI) against the C oracle on the same batch."""
import numpy as np

import oracle_c
from excord_lr_b200 import api
from excord_lr_b200.batch import EVENT_DTYPE, ExlrParams, HostBatch


def gpu_available() -> bool:
    try:
        return api.load_library().exlr_device_count() > 0
    except Exception:
        return False


def describe_read(hb: HostBatch, r: int) -> str:
    co, so = hb.cigar_off, hb.sa_off
    ops = hb.cigar[int(co[r]):int(co[r + 1])]
    cig = "".join("%d%s" % (v >> 4, "MIDNSHP=X??????!"[v & 15]) for v in ops[:60].tolist())
    sa = hb.sa_bytes[int(so[r]):int(so[r + 1])].tobytes()[:300]
    return (f"read {r}: tid={hb.tid[r]} pos={hb.pos[r]} flag={hb.flag[r]} mapq={hb.mapq[r]} sa_kind={hb.sa_kind[r]} "
            f"n_ops={len(ops)} cigar={cig}{'...' if len(ops) > 60 else ''} sa={sa!r}")


def first_diff(a: np.ndarray, b: np.ndarray) -> int:
    n = min(len(a), len(b))
    neq = np.nonzero(a[:n] != b[:n])[0]
    return int(neq[0]) if len(neq) else n


def check_result(hb: HostBatch, p: ExlrParams, res, text: bytes, verbose=False, label=""):
    want = oracle_c.run(hb, p)
    assert res.status == want.status, f"{label} status gpu={res.status} (read {res.err_read}) oracle={want.status} (read {want.err_read})" + \
        (("\n" + describe_read(hb, min(res.err_read, want.err_read, hb.n_reads - 1))) if hb.n_reads else "")
    if want.status != 0:
        assert res.err_read == want.err_read, f"{label} err_read gpu={res.err_read} oracle={want.err_read}"
        assert res.n_err_lines == want.n_err_lines, f"{label} n_err_lines gpu={res.n_err_lines} oracle={want.n_err_lines}"
        got_ev = res.events[:res.n_valid_lines()]
        want_ev = want.events                           # the oracle keeps exactly the lines the reference had written
    else:
        got_ev, want_ev = res.events, want.events
        assert res.n_kept == want.n_kept, f"{label} n_kept {res.n_kept} != {want.n_kept}"
        assert res.n_sa_reads == want.n_sa_reads, f"{label} n_sa_reads {res.n_sa_reads} != {want.n_sa_reads}"
        assert res.n_cap_dropped == want.n_cap_dropped, f"{label} n_cap_dropped {res.n_cap_dropped} != {want.n_cap_dropped}"
        if not np.array_equal(res.line_off, want.line_off):
            d = first_diff(res.line_off, want.line_off)
            r = max(0, d - 1)
            raise AssertionError(f"{label} line_off differs at {d}: gpu={res.line_off[d]} oracle={want.line_off[d]}\n" + describe_read(hb, r))
    if len(got_ev) != len(want_ev) or got_ev.tobytes() != want_ev.tobytes():
        d = first_diff(got_ev, want_ev)
        g = got_ev[d] if d < len(got_ev) else None
        w = want_ev[d] if d < len(want_ev) else None
        r = int(w["read_idx"]) if w is not None else (int(g["read_idx"]) if g is not None else 0)
        raise AssertionError(f"{label} events differ at {d} (gpu n={len(got_ev)}, oracle n={len(want_ev)}):\n gpu   ={g}\n oracle={w}\n" + describe_read(hb, r))
    want_text = oracle_c.format_lines(hb, want_ev, verbose)
    assert text == want_text, f"{label} formatted text differs"
    if getattr(res, "device_text", None) is not None:                  # kernels 5a/5b: the same bytes without the host formatter
        want_plain = want_text if not verbose else oracle_c.format_lines(hb, want_ev, False)
        assert res.device_text == want_plain, f"{label} device-formatted text differs"
    return want


def gpu_check(hb: HostBatch, p: ExlrParams, cigar_kernel=0, reads_per_cta=0, verbose=False, label="", max_events=0, long_records=0):
    res, text = api.extract(hb, p, 0, cigar_kernel, reads_per_cta, verbose, max_events, device_format=True, long_records=long_records)
    return check_result(hb, p, res, text, verbose, label), res
