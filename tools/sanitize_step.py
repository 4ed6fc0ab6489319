"""Small batches through every kernel path (all CIGAR strategies x long-record paths, device formatting), for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_step.py   (also --tool racecheck / initcheck)
(compute-sanitizer is closed on the pool this round was measured on; the same batches run as parity tests in tests/.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from excord_lr_b200 import api, synth
from excord_lr_b200.batch import ExlrParams
from randrec import rand_batch, rand_params

cases = [(synth.config(0, 0.1), ExlrParams.make(**synth.CONFIGS[0]["params"])), (synth.config(2, 0.002), ExlrParams.make(**synth.CONFIGS[2]["params"])),
         (synth.config(3, 0.002), ExlrParams.make(**synth.CONFIGS[3]["params"])), (rand_batch(5, 300), rand_params(5)), (rand_batch(6, 300), ExlrParams.make())]
n = 0
for hb, p in cases:
    for kernel, lr in ((0, 0), (1, 0), (2, 0), (3, 1), (3, 2), (3, 3)):
        res, text = api.extract(hb, p, 0, kernel, device_format=True, long_records=lr)
        n += 1
print("ran", n, "extractions")
