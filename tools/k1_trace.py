"""Per-CTA timeline of k1_flat (EXLR_OPT_TRACE = 1, flat block scan of every tile) for 1 and 3 waves of CTAs."""
import sys, ctypes as C
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from excord_lr_b200 import api, synth
from excord_lr_b200.batch import ExlrParams
hb=synth.config(1,1.0); p=ExlrParams.make(**synth.CONFIGS[1]['params'])
ex=api.Extractor(p,hb.ref_names,0); ex.set_option(api.EXLR_OPT_OVERLAP,0); ex.set_option(api.EXLR_OPT_TRACE,1)
ex.set_option(api.EXLR_OPT_CIGAR_KERNEL, api.CIGAR_KERNEL_FLAT)      # the per-CTA trace of k1_flat (flat block scan of every tile)
for waves in (1,3):
    ex.set_option(api.EXLR_OPT_K1_WAVES,waves)
    b=ex.batch_for(hb); b.upload()
    flush=torch.zeros(256<<20,dtype=torch.uint8,device='cuda')
    for i in range(4):
        flush.sum(); torch.cuda.synchronize(); b.submit_resident(); b.wait_resident()
    n=148*4*waves
    out=np.zeros(n*4,np.uint64)
    ex.lib.exlr_get_trace.argtypes=[C.c_void_p,C.c_void_p,C.c_uint32]
    ex.lib.exlr_get_trace(b.handle,out.ctypes.data,n)
    tr=out.reshape(n,4); tr=tr[tr[:,0]>0]
    t0=tr[:,0].min()
    start=(tr[:,0]-t0)/1e3; first=(tr[:,1]-tr[:,0])/1e3; dur=(tr[:,2]-tr[:,0])/1e3; end=(tr[:,2]-t0)/1e3
    tiles=(tr[:,3]&0xffffffff); scanned=(tr[:,3]>>np.uint64(32))
    print(f"waves {waves}: CTAs {len(tr)} kernel span {end.max():.1f} us; start p50 {np.median(start):.1f} p90 {np.percentile(start,90):.1f} max {start.max():.1f}; first-data latency p50 {np.median(first):.1f} p90 {np.percentile(first,90):.1f}; CTA duration p50 {np.median(dur):.1f} p90 {np.percentile(dur,90):.1f} max {dur.max():.1f}; end p50 {np.median(end):.1f} p90 {np.percentile(end,90):.1f}")
    print(f"   tiles/CTA {tiles.mean():.1f}, scanned tiles/CTA mean {scanned.mean():.2f} max {scanned.max()}; corr(dur, scanned) {np.corrcoef(dur, scanned.astype(float))[0,1]:.2f}; us per clean tile est {np.polyfit(scanned.astype(float), dur, 1)}")
    print('  ', b.timing().cigar_ms*1000)
    b.free()
