"""Multi-rank host logic on CPU: contig sharding plan + the count all-reduce over gloo (world_size 2).

No CUDA here: each rank fills its count tensor with the ORACLE's counts for its own contigs (test
infrastructure), so the test checks exactly what the N>1 path adds -- the partition and the collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nanomotif_b200 import sharding, synth
from oracle import restate as O

MOTIFS = [("GATC", 1), ("CC[AT]GG", 1), ("A", 0), ("GCAC......GTT", 2)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_bin(seed=5):
    rng = np.random.default_rng(seed)
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    for i, L in enumerate((9000, 20000, 4000, 15000, 7000)):
        seq = synth.random_sequence(rng, L, 0.5)
        name = f"c{i}"
        contigs[name] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=15, mod_types=("a",))
        cols["contig"].append(np.full(len(p["position"]), name, dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(np.where(p["strand"] == 0, "+", "-"))
        cols["fraction_mod"].append(p["fraction_mod"])
    return contigs, {k: np.concatenate(v) for k, v in cols.items()}


def _oracle_counts(contigs, pile):
    out = np.zeros((len(MOTIFS), 2), dtype=np.int64)
    for mi, (m, p) in enumerate(MOTIFS):
        out[mi] = O.motif_model_bin(pile["contig"], pile["position"], pile["strand"], pile["fraction_mod"], contigs, m, p,
                                    fast=True)
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        contigs, pile = _make_bin()
        owner = sharding.plan_shards([len(s) for s in contigs.values()], world)
        mine = sharding.local_contigs(contigs, owner, rank)
        counts = torch.from_numpy(_oracle_counts(mine, pile))
        sharding.allreduce_counts(counts)
        rows = sharding.gather_rows([(name, rank) for name in mine])
        q.put((rank, counts.numpy().tolist(), sorted(rows)))
    finally:
        dist.destroy_process_group()


def test_plan_shards_balances_and_keeps_bins():
    lengths = [100, 90, 80, 10, 10, 10, 5]
    owner = sharding.plan_shards(lengths, 2)
    loads = [sum(l for l, o in zip(lengths, owner) if o == r) for r in range(2)]
    assert abs(loads[0] - loads[1]) <= 40 and set(owner.tolist()) == {0, 1}  # LPT greedy: within one mid-sized contig
    assert sharding.plan_shards(lengths, 1).tolist() == [0] * 7
    # bins stay whole unless a bin exceeds 1/world of the total
    groups = [0, 0, 1, 1, 2, 2, 2]
    owner = sharding.plan_shards(lengths, 2, groups)
    assert owner[2] == owner[3] and owner[4] == owner[5] == owner[6]
    big = sharding.plan_shards([1000, 1000, 10], 2, [0, 0, 1])
    assert big[0] != big[1]


def test_allreduce_of_shard_counts_equals_unsharded_counts():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    contigs, pile = _make_bin()
    want = _oracle_counts(contigs, pile).tolist()
    for rank, counts, rows in results:
        assert counts == want, f"rank {rank}"
        assert [n for n, _ in rows] == sorted(contigs.keys())  # every contig owned by exactly one rank


def test_sharded_scorer_plan_is_pure_host_logic():
    """ShardedMultiBinScorer.plan: what every rank derives for itself before touching a device -- every contig owned
    once, bins whole unless one exceeds 1.25 / world of the assembly, lengths-only contigs accepted, same plan from
    sequences and from lengths."""
    bins = {"big": {f"big_{i}": "A" * n for i, n in enumerate((900, 700, 600, 400))},
            "s1": {"s1_0": "C" * 300, "s1_1": "C" * 80}, "s2": {"s2_0": "G" * 250}, "s3": {"s3_0": "T" * 90, "s3_1": "T" * 70}}
    owner, per_rank, split, bin_ranks = sharding.ShardedMultiBinScorer.plan(bins, 2)
    names = [n for cs in bins.values() for n in cs]
    assert len(owner) == len(names) and set(owner.tolist()) == {0, 1}
    got = sorted(n for r in per_rank for cs in r.values() for n in cs)
    assert got == sorted(names)                                    # every contig on exactly one rank
    assert split == {"big"} and bin_ranks["big"] == [0, 1]         # 2600 of 3390 bp > 1.25 / 2: split by contig
    for b in ("s1", "s2", "s3"):
        assert len(bin_ranks[b]) == 1                              # small bins stay whole
    loads = [sum(len(s) for cs in r.values() for s in cs.values()) for r in per_rank]
    assert abs(loads[0] - loads[1]) <= 400
    by_len = {b: {n: len(s) for n, s in cs.items()} for b, cs in bins.items()}
    owner2, per_rank2, split2, _ = sharding.ShardedMultiBinScorer.plan(by_len, 2)
    assert owner2.tolist() == owner.tolist() and split2 == split   # lengths are all a rank needs of foreign contigs
    assert [{b: list(cs) for b, cs in r.items()} for r in per_rank2] == [{b: list(cs) for b, cs in r.items()} for r in per_rank]
    # as many equal bins as ranks: the 25 % margin keeps them whole (cfg3's 8-bin sample on 8 GPUs)
    eight = {f"b{i}": {f"b{i}_c{j}": 1000 + 3 * i + j for j in range(5)} for i in range(8)}
    _, per8, split8, _ = sharding.ShardedMultiBinScorer.plan(eight, 8)
    assert not split8 and all(len(r) == 1 for r in per8)
    assert sharding.ShardedMultiBinScorer.plan(bins, 1)[1][0].keys() == bins.keys()


def _plan_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        contigs, pile = _make_bin()
        bins = {"bin_a": {k: contigs[k] for k in ("c0", "c1")}, "bin_b": {k: contigs[k] for k in ("c2", "c3", "c4")}}
        _, per_rank, split, _ = sharding.ShardedMultiBinScorer.plan(bins, world)
        # what submit() does on a device, with the oracle standing in for the scan: rows of requests whose bin is not
        # local stay zero, ONE all-reduce of the stacked tensor replicates every request's counts on every rank
        requests = [("bin_a", MOTIFS[:2]), ("bin_b", MOTIFS), ("bin_a", MOTIFS[2:])]
        rows = []
        for b, motifs in requests:
            local = per_rank[rank].get(b, {})
            for m, p in motifs:
                rows.append(O.motif_model_bin(pile["contig"], pile["position"], pile["strand"], pile["fraction_mod"], local, m, p,
                                              fast=True) if local else (0, 0))
        t = torch.tensor(rows, dtype=torch.int64)
        sharding.allreduce_counts(t)
        q.put((rank, t.tolist(), sorted(split)))
    finally:
        dist.destroy_process_group()


def test_stacked_request_tensor_all_reduce_replicates_counts():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_plan_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    contigs, pile = _make_bin()
    bins = {"bin_a": {k: contigs[k] for k in ("c0", "c1")}, "bin_b": {k: contigs[k] for k in ("c2", "c3", "c4")}}
    want = []
    for b, motifs in [("bin_a", MOTIFS[:2]), ("bin_b", MOTIFS), ("bin_a", MOTIFS[2:])]:
        for m, p in motifs:
            want.append(list(O.motif_model_bin(pile["contig"], pile["position"], pile["strand"], pile["fraction_mod"], bins[b], m, p,
                                               fast=True)))
    for rank, got, split in results:
        assert got == want, f"rank {rank}"


# ---------------------------------------------------------------------------------------------
# a contig that alone exceeds a rank's share is cut into position ranges (cfg 2 on several GPUs)
# ---------------------------------------------------------------------------------------------
LONG_MOTIFS = MOTIFS + [("A" + "." * 38 + "T", 0), ("T" + "." * 38 + "A", 39), ("G[AG].GAAG[CT]", 5), ("A" + "." * 60 + "C", 0)]


def _huge_bin(seed=11, length=6000, extra=(700, 300)):
    rng = np.random.default_rng(seed)
    bins, cols = {"mono": {}, "small": {}}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    for b, name, L in [("mono", "chrom", length)] + [("small", f"s{i}", n) for i, n in enumerate(extra)]:
        seq = synth.random_sequence(rng, L, 0.5, 2e-3 if name == "chrom" else 0.0)  # N runs in the big contig
        bins[b][name] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=15, mod_types=("a",))
        cols["contig"].append(np.full(len(p["position"]), name, dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(np.where(p["strand"] == 0, "+", "-"))
        cols["fraction_mod"].append(p["fraction_mod"])
    return bins, {k: np.concatenate(v) for k, v in cols.items()}


def _rank_counts(per_rank_bins, pile, motifs):
    """What one rank's MultiBinScorer(pieces=...) computes, with the oracle standing in for the scan: the rank's texts
    as its contigs, the rows moved to the pieces by sharding.remap_split_rows (torch, here on the CPU)."""
    texts, pieces = {}, []
    for cs in per_rank_bins.values():
        for name, seq in cs.items():
            if isinstance(seq, sharding.ContigPiece):
                texts[name] = seq.text
                pieces.append((name, seq.name, seq.a, seq.b, seq.shift))
            else:
                texts[name] = seq
    index = {n: i for i, n in enumerate(texts)}
    cid = torch.tensor([index.get(n, -1) for n in pile["contig"]], dtype=torch.int32)
    pos = torch.from_numpy(np.asarray(pile["position"], dtype=np.int64))
    if pieces:
        cid, pos = sharding.remap_split_rows(cid, pos, [(index[c], index[o], a, b, s) for o, c, a, b, s in pieces])
    names = np.array(list(texts) + ["?"], dtype=object)[cid.numpy()]
    out = np.zeros((len(motifs), 2), dtype=np.int64)
    for mi, (m, p) in enumerate(motifs):
        out[mi] = O.motif_model_bin(names, pos.numpy(), pile["strand"], pile["fraction_mod"], texts, m, p, fast=True)
    return out


@pytest.mark.parametrize("world,length", [(2, 6000), (4, 6000), (8, 20000), (3, 5000)])
def test_counts_of_contig_pieces_add_up_to_the_contig(world, length):
    bins, pile = _huge_bin(length=length)
    owner, per_rank, split, bin_ranks = sharding.ShardedMultiBinScorer.plan(bins, world)
    assert owner[0] == -1 and "mono" in split and len(bin_ranks["mono"]) == min(world, len(bin_ranks["mono"])) > 1
    pieces = sorted((p.a, p.b) for r in per_rank for cs in r.values() for p in cs.values() if isinstance(p, sharding.ContigPiece))
    assert pieces[0][0] == 0 and pieces[-1][1] == length
    assert all(x[1] == y[0] for x, y in zip(pieces, pieces[1:]))  # the ranges tile the contig
    for r in per_rank:  # local names are unique, the first piece of a contig on a rank keeps the contig's name
        mine = [n for cs in r.values() for n, p in cs.items() if isinstance(p, sharding.ContigPiece)]
        assert len(set(mine)) == len(mine) and (not mine or "chrom" in mine)
    want = _oracle_counts_motifs(bins, pile, LONG_MOTIFS)
    got = sum(_rank_counts(r, pile, LONG_MOTIFS) for r in per_rank)
    assert got.tolist() == want.tolist()
    assert want[:, 0].sum() > 0 and want[-1].sum() > 0  # the 62-position motif has joined rows
    # without splitting the big contig sits on one rank
    owner0, per0, _, _ = sharding.ShardedMultiBinScorer.plan(bins, world, split_contigs=False)
    assert owner0[0] >= 0 and not any(isinstance(p, sharding.ContigPiece) for r in per0 for cs in r.values() for p in cs.values())


def _oracle_counts_motifs(bins, pile, motifs):
    contigs = {n: s for cs in bins.values() for n, s in cs.items()}
    out = np.zeros((len(motifs), 2), dtype=np.int64)
    for mi, (m, p) in enumerate(motifs):
        out[mi] = O.motif_model_bin(pile["contig"], pile["position"], pile["strand"], pile["fraction_mod"], contigs, m, p, fast=True)
    return out


def test_two_pieces_of_one_contig_on_one_rank_and_lengths_only():
    """Pieces of a contig that land on one rank without being neighbours get their own local names; a rank needs the
    text of its own pieces only."""
    a, b = sharding.ContigPiece("x", 0, 1024, 5000, "ACGT" * 1250), sharding.ContigPiece("x", 2048, 3072, 5000, "ACGT" * 1250)
    assert (a.lo, a.hi, a.shift, len(a)) == (0, 1024 + 64, 0, 1088) and (b.lo, b.hi, b.shift) == (2048 - 64, 3072 + 64, 1984)
    assert a.text == ("ACGT" * 1250)[:1088] and b.text == ("ACGT" * 1250)[1984:3136]
    with pytest.raises(ValueError, match="length only"):
        sharding.ContigPiece("x", 0, 1024, 5000, 5000).text
    cid = torch.tensor([0, 0, 0, 0, 1, 0], dtype=torch.int32)
    pos = torch.tensor([5, 1024, 2048, 3071, 7, 4999])
    ncid, npos = sharding.remap_split_rows(cid, pos, [(0, 0, 0, 1024, 0), (0, 2, 2048, 3072, 1984)])
    assert ncid.tolist() == [0, -1, 2, 2, 1, -1] and npos.tolist() == [5, 1024, 64, 1087, 7, 4999]
    assert cid.tolist() == [0, 0, 0, 0, 1, 0]  # inputs untouched
    from nanomotif_b200 import _lib

    assert sharding.HALO_BP >= _lib.MAX_MOTIF_LEN  # an occurrence reaches at most MAX_MOTIF_LEN - 1 bp past its modified base
    assert sharding.split_ranges(5000, 3) == [(0, 1536), (1536, 3584), (3584, 5000)]
    assert sharding.split_ranges(1500, 4) == [(0, 1500)]  # too short to cut
    # contigs of other ranks by length: the same plan as from sequences
    bins, _ = _huge_bin()
    by_len = {b: {n: len(s) for n, s in cs.items()} for b, cs in bins.items()}
    p1, p2 = sharding.ShardedMultiBinScorer.plan(bins, 4), sharding.ShardedMultiBinScorer.plan(by_len, 4)
    assert p1[0].tolist() == p2[0].tolist() and p1[2] == p2[2] and p1[3] == p2[3]
    assert [[(b, n) for b, cs in r.items() for n in cs] for r in p1[1]] == [[(b, n) for b, cs in r.items() for n in cs] for r in p2[1]]
