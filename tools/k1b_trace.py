"""Per-warp timeline of kernel 1b (EXLR_OPT_TRACE): where a flagged step's warp spends its time."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from excord_lr_b200 import api, synth
from excord_lr_b200.batch import ExlrParams
hb = synth.config(1, 1.0); p = ExlrParams.make(**synth.CONFIGS[1]['params'])
ex = api.Extractor(p, hb.ref_names, 0); ex.set_option(api.EXLR_OPT_OVERLAP, 0); ex.set_option(7, 1)
b = ex.batch_for(hb); b.upload()
flush = torch.zeros(256 << 20, dtype=torch.uint8, device='cuda')
for i in range(4):
    flush.sum(); torch.cuda.synchronize(); b.submit_resident(); b.wait_resident()
n = 8192
out = np.zeros(n * 4, np.uint64)
ex.lib.exlr_get_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
ex.lib.exlr_get_trace(b.handle, out.ctypes.data, n)
tr = out.reshape(n, 4); tr = tr[tr[:, 0] > 0]
t0 = tr[:, 0].min()
start = (tr[:, 0] - t0) / 1e3; search = (tr[:, 1] - tr[:, 0]) / 1e3
walk = (tr[:, 2] - tr[:, 1]) / 1e3
total = (tr[:, 3] - tr[:, 0]) / 1e3
q = lambda x: "p10 %.1f p50 %.1f p90 %.1f max %.1f" % (np.percentile(x, 10), np.median(x), np.percentile(x, 90), x.max())
print("flagged steps traced", len(tr), "cigar path", b.timing().cigar_ms * 1e3, "screen", b.timing().screen_ms * 1e3)
print("step start after first start:", q(start))
print("list entry + search:", q(search)); print("offsets+claim+walk (first 32 records):", q(walk)); print("whole step:", q(total))
print("span (last end):", (start + total).max())
