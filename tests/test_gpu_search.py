"""The motif search driven by the GPU operators reproduces the trace of the REAL reference search loop, and
the lock-step multi-bin driver (one scan launch per step for all bins) gives the same result as bin-by-bin."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from search_common import build_inputs, check_against_trace, load_trace


def _setup(nmb, contigs, pile, spec, scorer):
    from nanomotif_b200 import growth, search

    asm = scorer.owner.assembly if hasattr(scorer, "owner") else scorer.assembly
    ids = np.array([asm.index[c] for c in contigs])
    cid = np.full(len(pile["position"]), -1, dtype=np.int64)
    for name, i in zip(contigs, ids):
        cid[pile["contig"] == name] = i
    strand = (pile["strand"] == "-").astype(np.uint8)
    sel = cid >= 0
    # windows in the reference's order: per contig '+' then '-' (growth.methylation_windows iterates assembly order)
    sub = {k: v[sel] for k, v in pile.items()}
    local = {c: i for i, c in enumerate(contigs)}
    windows = _windows_for(growth, asm, contigs, sub, strand[sel], spec)
    random.seed(spec["random_seed"])
    bin_pssm = growth.background_pssm(asm, contigs, search.CANONICAL[spec["mod_type"]], spec["padding"])
    return search.GpuBinBackend(scorer, windows), bin_pssm, windows.shape[0]


def _windows_for(growth, asm, contigs, pile, strand, spec):
    pad, high = spec["padding"], spec["high"]
    ci, pos, st = [], [], []
    conf = pile["fraction_mod"] >= high
    for name, seq in contigs.items():
        for s in (0, 1):
            p = pile["position"][conf & (pile["contig"] == name) & (strand == s)]
            p = p[(p > pad) & (p < len(seq) - pad)]
            ci.append(np.full(len(p), asm.index[name]))
            pos.append(p)
            st.append(np.full(len(p), s, dtype=np.uint8))
    return growth.DeviceDNAarray.from_positions(asm, np.concatenate(ci), np.concatenate(pos), np.concatenate(st), pad)


@pytest.fixture(scope="module")
def nmb():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nanomotif_b200

    return nanomotif_b200


def test_gpu_search_reproduces_reference_trace(nmb):
    from nanomotif_b200 import search

    trace = load_trace()
    spec = trace["spec"]
    contigs, pile = build_inputs(spec)
    scorer = nmb.BinScorer(pile, contigs, spec["low"], spec["high"])
    backend, bin_pssm, total = _setup(nmb, contigs, pile, spec, scorer)
    assert total == trace["total_windows"]
    np.testing.assert_array_equal(bin_pssm, np.array(trace["bin_pssm"]))
    rounds = []
    co = search.find_candidates(spec["mod_type"], spec["padding"], bin_pssm, total, min_kl=spec["min_kl"],
                                score_threshold=spec["score_threshold"], trace=rounds)
    check_against_trace(trace, search.run(co, backend), rounds)


def test_cfg1_bundled_assembly_on_gpu(nmb):
    """BASELINE.json configs[0] on the GPU operators: the bundled geobacillus plasmids with the shipped golden motifs
    planted; every node's counts / scores equal the trace recorded from the REAL reference search (whose scoring calls
    were answered by the oracle), i.e. every visited motif's counts equal the oracle's."""
    from nanomotif_b200 import search

    trace = load_trace("search_trace_cfg1.json")
    spec = trace["spec"]
    contigs, pile = build_inputs(spec)
    scorer = nmb.BinScorer(pile, contigs, spec["low"], spec["high"])
    backend, bin_pssm, total = _setup(nmb, contigs, pile, spec, scorer)
    assert total == trace["total_windows"]
    np.testing.assert_array_equal(bin_pssm, np.array(trace["bin_pssm"]))
    rounds = []
    co = search.find_candidates(spec["mod_type"], spec["padding"], bin_pssm, total, min_kl=spec["min_kl"],
                                score_threshold=spec["score_threshold"], trace=rounds)
    graph, best = search.run(co, backend)
    check_against_trace(trace, (graph, best), rounds)
    assert [m.new_stripped_motif().string for m in best[:3]] == ["GATC", "ACCCA", "CCAAAT"]


def test_lockstep_multibin_search(nmb):
    from nanomotif_b200 import search

    trace = load_trace()
    spec = trace["spec"]
    bins, piles = {}, []
    for b, seed in enumerate((spec["seed"], 77, 78)):
        s = dict(spec, seed=seed, planted=spec["planted"] if b == 0 else [["CTGCAG", 4, "a"], ["GA[AG]TC", 1, "a"]])
        contigs, pile = build_inputs(s)
        contigs = {f"bin{b}_{k}": v for k, v in contigs.items()}
        pile["contig"] = np.array([f"bin{b}_{c}" for c in pile["contig"]], dtype=object)
        pile["mod_type"] = np.full(len(pile["position"]), "a", dtype=object)
        bins[f"bin{b}"] = contigs
        piles.append(pile)
    pile = {k: np.concatenate([p[k] for p in piles]) for k in piles[0]}
    multi = nmb.MultiBinScorer(pile, bins, ["a"], spec["low"], spec["high"])

    def make(b):
        ctx = multi.context(f"bin{b}", "a")
        backend, bin_pssm, total = _setup(nmb, bins[f"bin{b}"], pile, spec, ctx)
        t = []
        co = search.find_candidates("a", spec["padding"], bin_pssm, total, min_kl=spec["min_kl"],
                                    score_threshold=spec["score_threshold"], trace=t)
        return co, backend, t

    from nanomotif_b200 import growth

    searches = [make(b) for b in range(3)]
    # all windows in one pool: score, expand and remove requests of the three searches are each ONE launch
    pool = growth.WindowPool([be.windows for _, be, _ in searches])
    pooled = [(co, search.PoolBackend(be.scorer, pool, slot)) for slot, (co, be, _) in enumerate(searches)]
    results = search.run_lockstep(pooled, search.gpu_batch_score, search.gpu_batch_expand, search.gpu_batch_remove)
    check_against_trace(trace, results[0], searches[0][2])  # bin 0 is the golden input
    for b in (1, 2):  # the others equal their own sequential runs
        co, be, t = make(b)
        seq_graph, seq_best = search.run(co, be)
        graph, best = results[b]
        assert [m.string for m in best] == [m.string for m in seq_best] and len(best) >= 1
        assert {m.string: (d["model"]._alpha, d["model"]._beta) for m, d in graph.nodes.items()} == \
               {m.string: (d["model"]._alpha, d["model"]._beta) for m, d in seq_graph.nodes.items()}
        assert searches[b][2] == t


def test_merge_groups_on_gpu_reproduce_reference_decisions(nmb):
    """The merge re-scoring step (find_motifs_bin.py:1438-1533) of three (bin, mod_type) groups in lock-step: the two
    score requests of every group share one K2 launch each; bin 0 is the golden input recorded from the reference."""
    from nanomotif_b200 import search
    from nanomotif_b200.model import BetaBernoulliModel
    from search_common import check_merge_against_trace, load_merge_trace, merge_inputs

    golden = load_merge_trace()
    spec = load_trace()["spec"]
    bins, piles = {}, []
    for b, seed in enumerate((spec["seed"], 77)):
        contigs, pile = build_inputs(dict(spec, seed=seed))
        contigs = {f"bin{b}_{k}": v for k, v in contigs.items()}
        pile["contig"] = np.array([f"bin{b}_{c}" for c in pile["contig"]], dtype=object)
        pile["mod_type"] = np.full(len(pile["position"]), "a", dtype=object)
        bins[f"bin{b}"] = contigs
        piles.append(pile)
    pile = {k: np.concatenate([p[k] for p in piles]) for k in piles[0]}
    multi = nmb.MultiBinScorer(pile, bins, ["a"], spec["low"], spec["high"])
    launches = []
    real = multi.score_batch
    multi.score_batch = lambda reqs: (launches.append(len(reqs)), real(reqs))[1]
    for run in golden["runs"]:
        groups, traces = [], []
        for b in (0, 1, 0):
            rows, clusters = merge_inputs(golden, run, search.Motif, BetaBernoulliModel)
            t = []
            traces.append(t)
            backend = search.PoolBackend(multi.context(f"bin{b}", "a"), None, 0)
            groups.append((search.merge_group(rows, clusters, run["merge_threshold"], trace=t), backend))
        launches.clear()
        results = search.run_lockstep(groups, search.gpu_batch_score)
        assert launches == [3, 3]  # two launches for all groups together
        for g in (0, 2):
            check_merge_against_trace(run, results[g], traces[g])
        assert len(results[1]) >= 1
