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
