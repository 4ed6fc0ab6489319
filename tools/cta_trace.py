"""Per-CTA timelines of the small kernels (EXLR_OPT_TRACE = 2..6): start skew, time to the mid mark, CTA duration, kernel span."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from excord_lr_b200 import api, synth
from excord_lr_b200.batch import ExlrParams
hb = synth.config(1, 1.0); p = ExlrParams.make(**synth.CONFIGS[1]['params'])
ex = api.Extractor(p, hb.ref_names, 0); ex.set_option(api.EXLR_OPT_OVERLAP, 0)
ex.lib.exlr_get_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
b = ex.batch_for(hb); b.upload()
flush = torch.zeros(256 << 20, dtype=torch.uint8, device='cuda')
names = {2: ("k0_classify", "classify_ms", "before scan"), 3: ("k3a_sa_cigar", "sa_cigar_ms", "n_sa loaded"), 4: ("k3b_sa_events", "sa_parse_ms", "before phase 3"),
         5: ("k4a_line_scan", "scan_ms", "before scan"), 6: ("k4b_place", "place_ms", "counts loaded")}
q = lambda x: "p10 %.1f p50 %.1f p90 %.1f max %.1f" % (np.percentile(x, 10), np.median(x), np.percentile(x, 90), x.max())
for sel, (name, key, midname) in names.items():
    ex.set_option(7, sel)
    for i in range(3):
        flush.sum(); torch.cuda.synchronize(); b.submit_resident(); b.wait_resident()
    out = np.zeros(8192 * 4, np.uint64)
    ex.lib.exlr_get_trace(b.handle, out.ctypes.data, 8192)
    tr = out.reshape(8192, 4); tr = tr[tr[:, 0] > 0]
    t0 = tr[:, 0].min()
    start = (tr[:, 0] - t0) / 1e3; mid = (tr[:, 1].astype(np.int64) - tr[:, 0].astype(np.int64)) / 1e3; dur = (tr[:, 2] - tr[:, 0]) / 1e3
    print(f"{name}: stage {getattr(b.timing(), key) * 1e3:.1f} us, CTAs traced {len(tr)}, span {(start + dur).max():.1f} us")
    print(f"    start {q(start)} | {midname}: {q(mid[mid > 0]) if (mid > 0).any() else '-'} | CTA duration {q(dur)}")
