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
    (bin id per contig) a bin's contigs stay together unless the bin alone exceeds 1.25 x 1/world_size of
    the total, in which case its contigs are spread individually (then the all-reduce merges the bin).  The
    margin keeps bins whole when there are about as many equal-sized bins as ranks."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    owner = np.zeros(n, dtype=np.int32)
    if world_size <= 1 or n == 0:
        return owner
    if groups is None:
        units = [(int(lengths[i]), [i]) for i in range(n)]
    else:
        groups = np.asarray(groups)
        limit = 1.25 * lengths.sum() / world_size
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
        import torch

        self.scorer = BinScorer(table, mine, low_meth_threshold, high_meth_threshold, device) if mine else None
        self.device = self.scorer.assembly.device if mine else torch.device(
            "cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def score(self, motifs) -> np.ndarray:
        import torch

        motifs = list(motifs)
        if self.scorer is not None:
            c = self.scorer.counts_by_strand(motifs)
        else:  # a rank without contigs still takes part in the collective
            c = torch.zeros((len(motifs), 4), dtype=torch.int64, device=self.device)
        c = allreduce_counts(c).cpu().numpy()
        return np.stack([c[:, 0] + c[:, 2], c[:, 1] + c[:, 3]], axis=1)


def _contig_length(seq) -> int:
    """Length of a contig given as str, DNAsequence-like (.sequence) or -- for contigs another rank will own --
    as a plain int (every rank must know all lengths to derive the same shard plan, not all sequences)."""
    if isinstance(seq, (int, np.integer)):
        return int(seq)
    return len(seq if isinstance(seq, str) else seq.sequence)


class ShardedContext:
    """One (bin, mod_type) of a ShardedMultiBinScorer; `local` is the rank's BinContext or None."""

    def __init__(self, owner, bin_name, mod_type, local):
        self.owner, self.bin_name, self.mod_type, self.local = owner, bin_name, mod_type, local

    def score(self, motifs) -> np.ndarray:
        return self.owner.score_batch([(self, motifs)])[0]


class PendingScores:
    """Counts of one submitted batch: the scan is enqueued, the all-reduce runs asynchronously."""

    def __init__(self, requests, tensor, work):
        self.requests, self.tensor, self.work = requests, tensor, work

    def result(self) -> list:
        from .api import split_counts

        if self.work is not None:
            self.work.wait()  # orders the current stream after the collective
        return split_counts(self.tensor.cpu().numpy(), self.requests)


class ShardedMultiBinScorer:
    """The multi-GPU form of api.MultiBinScorer -- what replaces the reference's process pool over bins
    (nanomotif/find_motifs_bin.py:330-372) on a box with several GPUs.

    `plan_shards` keeps a bin's contigs on one rank unless the bin alone exceeds 1/world_size of the assembly; every
    rank packs only its own contigs and ingests only the pileup rows of those contigs.  The (replicated) search
    driver calls `score_batch` with the SAME requests on every rank: each rank scans the requests whose bin has local
    contigs in ONE launch and the stacked [total motifs, 4] count tensor is summed over the ranks with ONE all-reduce
    (NCCL over NVLink) -- that merges split bins and replicates the whole-bin results in the same step.
    `submit` returns before the collective has run, so the driver can enqueue the next batch while this one's counts
    are in flight.

    Give every rank the pileup rows of ITS bins: the partitioned form {(bin, mod_type): frame} (frames of bins
    without local contigs are skipped before any copy) or a table that already holds only the rank's rows.  Rows of
    other ranks' contigs in a shared table are ignored, but they are still copied to the device."""

    def __init__(self, pileup, bins: dict, mod_types, low_meth_threshold: float, high_meth_threshold: float,
                 rank: int, world_size: int, device=None, group=None):
        import torch

        from .api import MultiBinScorer

        self.rank, self.world_size, self.group = int(rank), int(world_size), group
        self.owner, per_rank, self.split_bins, self.bin_ranks = self.plan(bins, world_size)
        mine = per_rank[self.rank]
        for bin_name, local in mine.items():
            if any(isinstance(v, (int, np.integer)) for v in local.values()):
                raise ValueError(f"bin {bin_name}: this rank owns a contig that was given by length only")
        self.mod_types = list(mod_types)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        tables = pileup
        if isinstance(pileup, dict) and "position" not in pileup:  # partitioned {(bin, mod_type): frame}
            tables = {k: v for k, v in pileup.items() if not (isinstance(k, tuple) and k[0] in bins and k[0] not in mine)}
        self.local = MultiBinScorer(tables, mine, self.mod_types, low_meth_threshold, high_meth_threshold,
                                    self.device) if mine else None
        self._local_bins = set(mine)

    @staticmethod
    def plan(bins: dict, world_size: int):
        """The shard plan every rank derives for itself (pure host logic, no device): (owner rank of every contig in
        `bins` order, [{bin: {contig: sequence}} per rank], set of bins whose contigs ended up on several ranks,
        {bin: ranks holding part of it}).  Contigs may be given by length (int) instead of sequence."""
        lengths, groups = [], []
        for b, cs in enumerate(bins.values()):
            for seq in cs.values():
                lengths.append(_contig_length(seq))
                groups.append(b)
        owner = plan_shards(lengths, world_size, groups)
        per_rank = [dict() for _ in range(max(1, world_size))]
        split_bins, bin_ranks, at = set(), {}, 0
        for bin_name, cs in bins.items():
            own = owner[at:at + len(cs)]
            bin_ranks[bin_name] = sorted(set(own.tolist()))
            if len(bin_ranks[bin_name]) > 1:
                split_bins.add(bin_name)
            for (name, seq), r in zip(cs.items(), own.tolist()):
                per_rank[r].setdefault(bin_name, {})[name] = seq
            at += len(cs)
        return owner, per_rank, split_bins, bin_ranks

    def context(self, bin_name, mod_type) -> ShardedContext:
        if bin_name not in self.bin_ranks:
            raise KeyError(bin_name)
        local = self.local.context(bin_name, mod_type) if bin_name in self._local_bins else None
        return ShardedContext(self, bin_name, mod_type, local)

    def submit(self, requests) -> PendingScores:
        import torch
        import torch.distributed as dist

        requests = [(ctx, list(motifs)) for ctx, motifs in requests]
        local_requests = [(ctx.local, ms) for ctx, ms in requests]
        total = sum(len(ms) for _, ms in requests)
        if self.local is not None:
            t = self.local.score_batch_device(local_requests)
        else:  # a rank without contigs still takes part in the collective
            with torch.cuda.device(self.device):
                t = torch.zeros((total, 4), dtype=torch.int64, device=self.device)
        work = None
        if total and self.world_size > 1 and dist.is_available() and dist.is_initialized():
            work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return PendingScores(requests, t, work)

    def score_batch(self, requests) -> list:
        return self.submit(requests).result()
