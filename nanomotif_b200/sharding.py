"""Contig-sharded scoring across the GPUs of one box.

The reference parallelises over bins with a process pool and no communication
(nanomotif/find_motifs_bin.py:330-372).  Here one process drives one GPU (torchrun); whole contigs are
assigned to ranks by greedy bin-packing on base pairs, every rank scores the replicated motif list
against its own contigs, and the per-motif bin-level counts are summed with ONE all-reduce
(int64, sum) per scoring step -- NCCL over NVLink on GPUs, gloo in the CPU tests.  Per-contig outputs
(the contig x motif table) need no collective: rows are owned by the rank that owns the contig.
"""
from __future__ import annotations

import heapq
from typing import Mapping, Sequence

import numpy as np


def plan_shards(lengths: Sequence[int], world_size: int, groups: Sequence[int] | None = None) -> np.ndarray:
    """Rank of every contig.  Longest-first greedy packing balances total bp per rank.  With `groups`
    (bin id per contig) a bin's contigs stay together unless the bin alone exceeds 1/world_size of the
    total, in which case its contigs are spread individually (then the all-reduce merges the bin)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    owner = np.zeros(n, dtype=np.int32)
    if world_size <= 1 or n == 0:
        return owner
    if groups is None:
        units = [(int(lengths[i]), [i]) for i in range(n)]
    else:
        groups = np.asarray(groups)
        limit = lengths.sum() / world_size
        units = []
        for g in np.unique(groups):
            idx = np.flatnonzero(groups == g).tolist()
            total = int(lengths[idx].sum())
            if total > limit:
                units += [(int(lengths[i]), [i]) for i in idx]
            else:
                units.append((total, idx))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    for size, idx in sorted(units, key=lambda u: (-u[0], u[1][0])):
        load, r = heapq.heappop(heap)
        owner[idx] = r
        heapq.heappush(heap, (load + size, r))
    return owner


def local_contigs(contigs: Mapping[str, object], owner: np.ndarray, rank: int) -> dict:
    """The sub-dict of `contigs` (insertion order kept) owned by `rank`."""
    return {name: seq for i, (name, seq) in enumerate(contigs.items()) if owner[i] == rank}


def allreduce_counts(counts, group=None):
    """Sum per-motif count tensors over the ranks in place (int64).  `counts` is a torch tensor on the
    rank's device (CUDA -> NCCL) or on the CPU (gloo).  No-op without an initialised process group."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def gather_rows(rows, group=None):
    """Concatenate per-contig result rows of all ranks on every rank (all_gather_object; rows are small
    host objects such as the contig x motif table slices)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(rows)
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, list(rows), group=group)
    return [r for part in out for r in part]


class ShardedBinScorer:
    """BinScorer over the rank's share of a bin's contigs + all-reduce of the counts.

    Every rank must call `score` with the same motif list (the search is replicated; only the scan is
    sharded).  Requires an initialised torch.distributed process group when world_size > 1."""

    def __init__(self, pileup, contigs, low_meth_threshold, high_meth_threshold, rank: int, world_size: int,
                 device=None):
        from .api import BinScorer
        from .pileup import PileupTable

        names = list(contigs.keys())
        lengths = [len(c if isinstance(c, str) else c.sequence) for c in contigs.values()]
        self.owner = plan_shards(lengths, world_size)
        self.rank, self.world_size = rank, world_size
        mine = local_contigs(contigs, self.owner, rank)
        table = PileupTable.from_frame(pileup)
        if table.contig is not None and len(mine) < len(names):
            keep = np.isin(np.asarray(table.contig).astype(str), np.array(list(mine.keys()), dtype=str))
            table = table.take(keep)
        self.scorer = BinScorer(table, mine, low_meth_threshold, high_meth_threshold, device) if mine else None

    def score(self, motifs) -> np.ndarray:
        import torch

        motifs = list(motifs)
        if self.scorer is not None:
            c = self.scorer.counts_by_strand(motifs)
        else:  # a rank without contigs still takes part in the collective
            c = torch.zeros((len(motifs), 4), dtype=torch.int64, device="cuda")
        c = allreduce_counts(c).cpu().numpy()
        return np.stack([c[:, 0] + c[:, 2], c[:, 1] + c[:, 3]], axis=1)
