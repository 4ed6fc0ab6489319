#!/usr/bin/env python
"""SASS opcode histogram per kernel of libexlr_cuda.so (cuobjdump -sass), written to profiles/.  Shows at a glance what each
kernel is made of: LDG.E.128 / PRMT streaming in k1a, UBLKCP + SYNCS (TMA bulk copy + mbarrier) in k1_flat, shared-memory
traffic and branches in k3b, and that nothing on this path issues tensor-core (HMMA / UTCMMA) instructions.
  python tools/sass_hist.py > profiles/r2_sass_hist.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "excord_lr_b200", "csrc", "libexlr_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, hist, order, arch = None, {}, [], set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        if kern not in hist:
            hist[kern] = collections.Counter()
            order.append(kern)
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}   cubin arch: {', '.join(sorted(arch))}")
print("# per kernel: instruction count, then the opcodes by frequency (base mnemonic.modifiers as printed by cuobjdump)")
for k in order:
    h = hist[k]
    n = sum(h.values())
    grp = collections.Counter()
    for op, c in h.items():
        grp[op.split(".")[0]] += c
    tags = []
    for name, pat in (("128-bit global loads", r"^LDG\.E\.128"), ("TMA bulk copy", r"^UBLKCP"), ("mbarrier", r"^SYNCS"), ("PRMT", r"^PRMT"),
                      ("shared loads", r"^LDS"), ("shared stores", r"^STS"), ("local (spill/stack)", r"^(LDL|STL)"), ("atomics", r"^(ATOM|RED|ATOMG|ATOMS)"),
                      ("votes/shuffles", r"^(VOTE|SHFL|MATCH)"), ("tensor core", r"^(HMMA|IMMA|UTCMMA|UTCHMMA|QGMMA|HGMMA)"), ("f64", r"^D(ADD|MUL|FMA|SETP)|^MUFU\.RCP64")):
        c = sum(v for op, v in h.items() if re.search(pat, op))
        if c or name == "tensor core":
            tags.append(f"{name} {c}")
    print(f"\n{k}: {n} instructions   [{'; '.join(tags)}]")
    print("  " + "  ".join(f"{op} {c}" for op, c in grp.most_common(24)))
