"""Timeline of one device-resident step (EXLR_OPT_TRACE = 7): first CTA start and last CTA end of every kernel, relative to
the first kernel's start, with the CIGAR path beside the SA branch (as timed) and on one stream."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from excord_lr_b200 import api, synth
from excord_lr_b200.batch import ExlrParams
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
fmt = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # 1: with kernels 5a/5b (device-formatted lines)
quick = len(sys.argv) > 3                                   # any third argument: graph replay and direct launches of the default shape only
hb = synth.config(cfg, 1.0); p = ExlrParams.make(**synth.CONFIGS[cfg]['params'])
ex = api.Extractor(p, hb.ref_names, 0)
ex.set_option(api.EXLR_OPT_DEVICE_FORMAT, fmt)
ex.lib.exlr_get_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
b = ex.batch_for(hb); b.upload()
flush = torch.zeros(256 << 20, dtype=torch.uint8, device='cuda')
names = {2: "k0_classify", 8: "k1a_screen", 9: "k1b_claim", 10: "k1b_walk", 11: "k1c/k1d/k1_flat", 3: "k3a_sa_cigar", 4: "k3b_sa_events",
         5: "k4a_line_scan", 6: "k4b_place", 12: "k5a_line_bytes", 13: "k5b_format"}
ex.set_option(api.EXLR_OPT_STAGE_TIMING, 0)
ex.set_option(api.EXLR_OPT_TRACE, 7)
for overlap, k1a_ctas, graph in (((1, 8, 1), (2, 8, 1), (1, 8, 0), (2, 8, 0)) if quick else ((1, 8, 0), (1, 6, 0), (1, 4, 0), (0, 8, 0))):
    ex.set_option(api.EXLR_OPT_OVERLAP, overlap)
    ex.set_option(9, k1a_ctas)
    ex.set_option(api.EXLR_OPT_GRAPH, graph)
    for i in range(6):
        flush.sum(); torch.cuda.synchronize(); b.submit_resident(); b.wait_resident()
    out = np.zeros(16 * 4, np.uint64)
    ex.lib.exlr_get_trace(b.handle, out.ctypes.data, 16)
    tr = out.reshape(16, 4)
    rows = [(int(~tr[k, 0] & np.uint64(0xffffffffffffffff)), int(tr[k, 1]), int(tr[k, 3]), n) for k, n in names.items() if tr[k, 3] > 0]
    t0 = min(r[0] for r in rows)
    print(f"overlap={overlap} k1a CTAs/SM {k1a_ctas} graph={graph}: kernels_ms {b.timing().kernels_ms * 1e3:.1f} us (first kernel start -> last kernel end, CUDA events)")
    for s, e, n, name in sorted(rows):
        print(f"   {name:18s} CTAs {n:5d}  start {(s - t0) / 1e3:7.1f}  end {(e - t0) / 1e3:7.1f}  span {(e - s) / 1e3:6.1f} us")
