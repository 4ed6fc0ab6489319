"""Record sharding for N GPUs: no data-path collective, because every output line depends on exactly one record
(reference src/main.rs:204,211,609,769 re-initialise all state per iteration).

The input is cut into fixed-size batches in record order; batch `seq` goes to rank `seq % world` (round robin, SURVEY.md
§8e); each rank runs its batches on its own GPU; the outputs are re-assembled on the host by batch sequence number, which
restores the reference's output order (record order).
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Sequence, Tuple

from .batch import HostBatch


def plan_batches(n_reads: int, batch_reads: int) -> List[Tuple[int, int, int]]:
    """[(seq, first_record, last_record_exclusive)] covering 0..n_reads in order."""
    if batch_reads <= 0:
        raise ValueError("batch_reads must be positive")
    return [(s, a, min(a + batch_reads, n_reads)) for s, a in enumerate(range(0, n_reads, batch_reads))]


def rank_batches(plan: Sequence[Tuple[int, int, int]], rank: int, world: int) -> List[Tuple[int, int, int]]:
    return [b for b in plan if b[0] % world == rank]


def run_rank(hb: HostBatch, plan, rank: int, world: int, runner: Callable[[HostBatch], bytes]) -> List[Tuple[int, bytes]]:
    """Runs this rank's batches through `runner` (host batch -> formatted lines); returns [(seq, lines)]."""
    return [(seq, runner(hb.slice(a, b))) for seq, a, b in rank_batches(plan, rank, world)]


def merge_ordered(parts: Iterable[Sequence[Tuple[int, bytes]]]) -> bytes:
    """Concatenate per-rank [(seq, lines)] lists in batch order."""
    allp = sorted((p for part in parts for p in part), key=lambda x: x[0])
    seqs = [s for s, _ in allp]
    if seqs != list(range(len(seqs))):
        raise ValueError("missing or duplicated batch in the gathered results")
    return b"".join(t for _, t in allp)
