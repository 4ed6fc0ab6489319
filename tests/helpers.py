"""Shared test helpers: conversions between the SoA batch and the text-level Python oracle."""
import numpy as np

import oracle as pyoracle
from excord_lr_b200.batch import BAM_OPS, SA_NONE, SA_STRING, ExlrParams, HostBatch


def py_params(p: ExlrParams, verbose=False) -> "pyoracle.Params":
    return pyoracle.Params(mapq=p.mapq, exclude_flag=p.exclude_flag, exclude_secondary=bool(p.exclude_secondary),
                           exclude_unmapped=bool(p.exclude_unmapped), indel_min=p.indel_min, merge_min=p.merge_min,
                           ins_clip_min=p.ins_clip_min, split_only=bool(p.split_only),
                           max_pct_overlap=p.max_pct_overlap, max_supp_alignm=p.max_supp_alignm, verbose=verbose)


def py_records(hb: HostBatch):
    """HostBatch -> list of oracle.Record (text level)."""
    recs = []
    qn = hb.qnames if hb.qnames is not None else ["r%09d" % i for i in range(hb.n_reads)]
    cig = hb.cigar.tolist()
    co = hb.cigar_off.tolist()
    so = hb.sa_off.tolist()
    sab = hb.sa_bytes.tobytes()
    for i in range(hb.n_reads):
        ops = [(v & 0xF, v >> 4) for v in cig[co[i]:co[i + 1]]]
        t = int(hb.tid[i])
        contig = hb.ref_names[t] if 0 <= t < len(hb.ref_names) else None
        k = int(hb.sa_kind[i])
        sa = None if k == SA_NONE else sab[so[i]:so[i + 1]].decode("latin-1")
        recs.append(pyoracle.Record(contig, int(hb.pos[i]), int(hb.flag[i]), int(hb.mapq[i]), ops, sa,
                                    sa_is_string=(k == SA_STRING), qname=qn[i]))
    return recs


def py_run(hb: HostBatch, p: ExlrParams, verbose=False):
    """-> (text, err_read or None)"""
    try:
        return pyoracle.run(py_records(hb), py_params(p, verbose)), None
    except pyoracle.ReferencePanic as e:
        return e.partial, e.read
