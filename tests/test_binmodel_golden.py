"""a3 / a4 / a8 and the loader filters against vectors recorded from the reference's OWN frame-taking functions (CPU).

motif_model_contig / motif_model_bin / get_parent_scores (find_motifs_bin.py:1265-1331, 1382-1433) and filter_pileup /
filter_pileup_minimummod_frequency (dataload.py:191-226) take polars frames; tests/golden/generate_binmodel_golden.py ran
them UNMODIFIED on numpy columns through oracle/minipolars.py (the handful of polars calls they make) and recorded the
results in tests/golden/binmodel_vectors.json.  Here the oracle's restatement of that glue is held against them; with
/root/reference mounted, the real functions are also run again, live, on fresh inputs -- including the unmodified
find_best_candidates, which must reproduce the committed search traces node by node."""
import hashlib
import importlib.util
import json
import os
import random
import tempfile

import numpy as np
import pytest

from oracle import ref_shim, restate as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(HERE, "golden", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    return spec, mod


@pytest.fixture(scope="module")
def G():
    with open(os.path.join(HERE, "golden", "binmodel_vectors.json")) as f:
        return json.load(f)


def build_inputs(spec):
    """The generator's inputs, regenerated from the seed (same code as tests/golden/generate_binmodel_golden.py)."""
    from nanomotif_b200 import synth

    rng = np.random.default_rng(spec["seed"])
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "mod_type", "fraction_mod", "Nvalid_cov")}
    for i, L in enumerate(spec["contig_lengths"]):
        seq = synth.random_sequence(rng, L, spec["gc"], spec["n_rate"])
        name = f"contig_{i}"
        contigs[name] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=spec["depth"], mod_types=tuple(spec["mod_types"]),
                               planted=[tuple(x) for x in spec["planted"]])
        n = len(p["position"])
        cols["contig"].append(np.full(n, name, dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(np.where(p["strand"] == 0, "+", "-").astype(object))
        cols["mod_type"].append(np.array(spec["mod_types"], dtype=object)[p["mod_type"]])
        cols["fraction_mod"].append(p["fraction_mod"])
        cols["Nvalid_cov"].append(p["Nvalid_cov"])
    return contigs, {k: np.concatenate(v) for k, v in cols.items()}


def digest(a) -> dict:
    a = np.ascontiguousarray(np.asarray(a), dtype=np.int64)
    return dict(n=int(a.size), sha1=hashlib.sha1(a.tobytes()).hexdigest(), head=a[:6].tolist())


@pytest.fixture(scope="module")
def inputs(G):
    contigs, pile = build_inputs(G["spec"])
    assert len(pile["position"]) == G["n_rows"] and int(pile["position"].sum()) == G["checksum"], \
        "the seeded generator no longer reproduces the inputs the vectors were recorded on"
    return contigs, pile


def test_oracle_bin_and_contig_counts_equal_the_reference_functions(G, inputs):
    contigs, pile = inputs
    for rec in G["bin"]:
        sel = pile["mod_type"] == rec["mod_type"]
        got = O.motif_model_bin(pile["contig"][sel], pile["position"][sel], pile["strand"][sel], pile["fraction_mod"][sel],
                                contigs, rec["motif"], rec["mod_pos"], rec["low"], rec["high"], fast=True)
        assert list(got) == rec["counts"], rec
    assert sum(r["counts"][0] for r in G["bin"]) > 1000
    for rec in G["contig"]:
        sel = (pile["mod_type"] == rec["mod_type"]) & (pile["contig"] == rec["contig"])
        for fast in (False, True):
            n_mod, n_nomod, pos = O.motif_model_contig(pile["position"][sel], pile["strand"][sel], pile["fraction_mod"][sel],
                                                       contigs[rec["contig"]], rec["motif"], rec["mod_pos"], fast=fast)
            assert [n_mod, n_nomod] == rec["counts"], rec
            assert {k: digest(v) for k, v in pos.items()} == rec["positions"], (rec["motif"], rec["contig"])


def test_oracle_parent_scores_equal_get_parent_scores(G, inputs):
    contigs, pile = inputs
    for rec in G["parents"]:
        sel = pile["mod_type"] == rec["mod_type"]
        cols = (pile["contig"][sel], pile["position"][sel], pile["strand"][sel], pile["fraction_mod"][sel])
        child = O.motif_model_bin(*cols, contigs, rec["motif"], rec["mod_pos"], fast=True)
        toks = O.split_motif(rec["motif"])
        want_positions = [i for i, t in enumerate(toks) if i != rec["mod_pos"] and t not in (".", "N")]
        assert [p["motif_position"] for p in rec["parents"]] == want_positions  # dict order = position order (:1411)
        for p in rec["parents"]:
            t = list(toks)
            t[p["motif_position"]] = "."
            assert "".join(t) == p["parent"] and p["mod_pos"] == rec["mod_pos"]
            parent = O.motif_model_bin(*cols, contigs, p["parent"], p["mod_pos"], fast=True)
            assert list(parent) == p["parent_counts"] and list(child) == p["child_counts"]
            score = O.predictive_evaluation_score(O.posterior(*child), O.posterior(*parent))
            assert score == pytest.approx(p["score"], rel=1e-12, abs=1e-12)


def test_oracle_filters_equal_the_reference_filters(G, inputs):
    _, pile = inputs
    rows = np.arange(len(pile["position"]))
    for rec in G["filters"]:
        if rec["kind"] == "coverage":
            keep = O.filter_pileup(pile["Nvalid_cov"], rec["min_coverage"])
        else:
            keep = O.filter_pileup_minimummod_frequency(pile["contig"], pile["mod_type"], pile["fraction_mod"],
                                                        rec["methylation_threshold"], rec["min_mod_frequency"],
                                                        rec["min_mods_pr_contig"])
            groups = sorted({f"{c}_{m}" for c, m in zip(pile["contig"][keep], pile["mod_type"][keep])})
            assert groups == rec["groups"]
            assert "contig_mod" not in rec["columns"]  # the helper column is dropped again (dataload.py:225)
        assert int(keep.sum()) == rec["n_kept"] and int(rows[keep].sum()) == rec["row_sum"], rec
    kinds = [r["n_kept"] for r in G["filters"] if r["kind"] == "min_mod_frequency"]
    assert 0 in kinds and max(kinds) > 0  # thresholds that keep nothing and thresholds that keep groups


# ---------------------------------------------------------------------------------------------
# live: the reference's functions through oracle/minipolars.py, wherever /root/reference is mounted
# ---------------------------------------------------------------------------------------------
needs_reference = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def nm():
    from oracle import minipolars

    ref = ref_shim.load_reference()
    minipolars.install(ref)
    return ref


@needs_reference
def test_live_reference_bin_model_and_filters(nm):
    from oracle import minipolars as mp

    spec = dict(seed=31337, contig_lengths=[9000, 3000, 150], gc=0.5, n_rate=5e-4,  # a bin that is not in the golden file
                depth=10, mod_types=["a", "m"], planted=[["GATC", 1, "a"]])
    contigs, pile = build_inputs(spec)
    frame = mp.DataFrame(pile)
    seqs = {k: nm.seq.DNAsequence(v) for k, v in contigs.items()}
    M, B = nm.motif.Motif, nm.model.BetaBernoulliModel
    for mt, motifs in (("a", [("GATC", 1), ("A", 0), ("..[AG]A[CT].", 3), ("A........T", 0)]), ("m", [("C", 0), ("CC[AT]GG", 1)])):
        sub = frame.filter(mp.col("mod_type") == mt)
        sel = pile["mod_type"] == mt
        for s, p in motifs:
            want = nm.find_motifs_bin.motif_model_bin(sub, seqs, M(s, p), B(), 0.3, 0.7).get_raw_counts()
            got = O.motif_model_bin(pile["contig"][sel], pile["position"][sel], pile["strand"][sel], pile["fraction_mod"][sel],
                                    contigs, s, p, fast=True)
            assert tuple(got) == tuple(want), (spec["seed"], mt, s)
    keep = O.filter_pileup_minimummod_frequency(pile["contig"], pile["mod_type"], pile["fraction_mod"], 0.7, 0.001, 20)
    ref = nm.dataload.filter_pileup_minimummod_frequency(mp.DataFrame(dict(pile, row=np.arange(frame.height))), 0.7, 0.001, 20)
    assert ref["row"].to_list() == np.flatnonzero(keep).tolist(), spec["seed"]


@needs_reference
@pytest.mark.parametrize("trace_name,spec_name", [("search_trace.json", "SPEC"), ("search_trace_cfg1.json", "CFG1_SPEC")])
def test_unmodified_find_best_candidates_reproduces_the_committed_trace(nm, trace_name, spec_name):
    """tests/golden/search_trace*.json were recorded with the candidate loop's polars glue RESTATED in the generator.
    Here the reference's find_best_candidates (find_motifs_bin.py:607-834) runs as it is -- real MotifSearcher, real
    motif_model_bin / get_parent_scores on a frame -- and must arrive at the same graph: every node's posterior, score,
    depth and visited flag, every edge, and the same candidates."""
    from oracle import minipolars as mp

    spec_, gsg = _load("generate_search_golden")
    import sys

    argv, sys.argv = sys.argv, ["generate_search_golden.py"]
    try:
        spec_.loader.exec_module(gsg)  # importing it only defines the specs and the input builder
    finally:
        sys.argv = argv
    spec = getattr(gsg, spec_name)
    with open(os.path.join(HERE, "golden", trace_name)) as f:
        trace = json.load(f)
    assert trace["spec"] == json.loads(json.dumps(spec))
    contigs, pile = gsg.build_inputs(spec)
    seqs = {k: nm.seq.DNAsequence(v) for k, v in contigs.items()}
    random.seed(spec["random_seed"])
    with tempfile.TemporaryDirectory() as out:
        graph, best = nm.find_motifs_bin.find_best_candidates(
            mp.DataFrame(dict(pile)), seqs, spec["mod_type"], "bin1", out, spec["low"], spec["high"], spec["padding"],
            min_kl=spec["min_kl"], score_threshold=spec["score_threshold"])
    assert [b.string for b in best] == trace["best_candidates"] + trace["missed"]
    nodes = sorted((n.string, int(d["model"]._alpha), int(d["model"]._beta), float(d["score"]), float(d["priority"]),
                    int(d["depth"]), bool(d["visited"])) for n, d in graph.nodes(data=True))
    want = sorted((n["motif"], n["alpha"], n["beta"], n["score"], n["priority"], n["depth"], n["visited"]) for n in trace["nodes"])
    assert len(nodes) == len(want)
    for a, b in zip(nodes, want):
        assert a[:3] == b[:3] and a[5:] == b[5:] and a[3] == pytest.approx(b[3], rel=1e-12, abs=1e-12) \
            and a[4] == pytest.approx(b[4], rel=1e-12, abs=1e-300), (a, b)
    assert sorted([u.string, v.string] for u, v in graph.edges()) == sorted(trace["edges"])


@needs_reference
def test_unmodified_merge_motifs_in_df_reproduces_the_committed_merge_trace(nm):
    """tests/golden/merge_trace.json was recorded with the loop body of merge_motifs_in_df restated without its frames.
    Here the reference function itself (find_motifs_bin.py:1436-1537: group_by, the real merge_motifs clustering, the
    real motif_model_bin / get_parent_scores on frames, concat) runs and must return the same rows, models and scores.
    Only the MotifSearchResult wrapper (a polars subclass that is not on the path) is bypassed."""
    import sys

    from oracle import minipolars as mp
    from search_common import build_inputs as search_inputs, load_merge_trace, load_trace

    golden, spec = load_merge_trace(), load_trace()["spec"]
    contigs, pile = search_inputs(spec)
    n = len(golden["motifs"])
    B = nm.model.BetaBernoulliModel
    motif_df = mp.DataFrame({"motif": [s for s, _ in golden["motifs"]], "score": [1.0] * n,
                             "mod_position": [p for _, p in golden["motifs"]], "reference": ["bin1"] * n, "mod_type": ["a"] * n,
                             "model": [B() for _ in range(n)]})
    pileup = mp.DataFrame(dict(pile, mod_type=np.full(len(pile["position"]), "a", dtype=object)))
    assembly = {k: nm.seq.DNAsequence(v) for k, v in contigs.items()}
    pl = sys.modules["polars"]
    saved = (pl.DataFrame, nm.motif.MotifSearchResult)
    pl.DataFrame, nm.motif.MotifSearchResult = mp.DataFrame, (lambda frame, *a, **k: frame)
    try:
        for run in golden["runs"]:
            res = nm.find_motifs_bin.merge_motifs_in_df(motif_df, pileup, assembly, {"bin1": list(contigs)}, spec["low"],
                                                        spec["high"], merge_threshold=run["merge_threshold"])
            # the order of the merged rows follows networkx's clique enumeration over a set of str-hashed nodes, i.e. the
            # interpreter's hash seed: compare the rows as a set
            got = {(m, int(p)): (model, score) for m, p, model, score in zip(
                res["motif"].to_list(), res["mod_position"].to_list(), res["model"].to_list(), res["score"].to_list())}
            assert len(got) == res.height == len(run["rows"])
            assert sorted(got) == sorted((r["motif"], r["mod_position"]) for r in run["rows"]), run["merge_threshold"]
            for want in run["rows"]:
                model, score = got[(want["motif"], want["mod_position"])]
                if want["merged"]:
                    assert [int(model._alpha), int(model._beta)] == want["model"]
                    assert float(score) == pytest.approx(want["score"], rel=1e-9, abs=1e-12)
            assert sum(r["merged"] for r in run["rows"]) > 0
    finally:
        pl.DataFrame, nm.motif.MotifSearchResult = saved


@needs_reference
@pytest.mark.parametrize("seed,mod_type,planted", [
    (4242, "m", [["CC[AT]GG", 1, "m"], ["GCGC", 1, "m"]]),
    (977, "a", [["GAGG", 1, "a"], ["CTA....TGC", 2, "a"], ["AC...GT", 0, "a"]]),
])
def test_search_driver_equals_the_unmodified_reference_search_on_new_inputs(nm, seed, mod_type, planted):
    """Beyond the two committed traces: a bin the repo has never seen (other seed, other motifs, 5mC as well as 6mA).
    Reference side: the unmodified find_best_candidates on a frame.  This side: search.find_candidates (the coroutine the
    lock-step GPU driver advances) answered by the oracle backend.  Same graph -- posteriors, scores, priorities, depth
    and visited flag of every node, same edges -- and the same candidates in the same order."""
    from oracle import minipolars as mp
    from nanomotif_b200 import search
    from search_common import OracleBackend, build_inputs as search_inputs, check_against_trace

    spec = dict(seed=seed, contig_lengths=[70000, 40000, 2500], gc=0.5, depth=15, mod_type=mod_type, planted=planted,
                padding=20, low=0.3, high=0.7, min_kl=0.05, score_threshold=1.5, random_seed=3)
    contigs, pile = search_inputs(spec)
    seqs = {k: nm.seq.DNAsequence(v) for k, v in contigs.items()}
    random.seed(spec["random_seed"])
    with tempfile.TemporaryDirectory() as out:
        res = nm.find_motifs_bin.find_best_candidates(
            mp.DataFrame(dict(pile)), seqs, mod_type, "bin1", out, spec["low"], spec["high"], spec["padding"],
            min_kl=spec["min_kl"], score_threshold=spec["score_threshold"])
    assert res is not None
    graph, best = res
    trace = dict(best_candidates=[b.string for b in best], missed=[],
                 nodes=[dict(motif=n.string, alpha=int(d["model"]._alpha), beta=int(d["model"]._beta), score=float(d["score"]),
                             priority=float(d["priority"]), depth=int(d["depth"]), visited=bool(d["visited"]))
                        for n, d in graph.nodes(data=True)],
                 edges=[[u.string, v.string] for u, v in graph.edges()])
    assert len(trace["nodes"]) > 10 and trace["best_candidates"]
    backend = OracleBackend(contigs, pile, spec)
    rounds = []
    co = search.find_candidates(mod_type, spec["padding"], backend.bin_pssm, backend.arr.shape[0], min_kl=spec["min_kl"],
                                score_threshold=spec["score_threshold"], trace=rounds)
    got_graph, got_best = search.run(co, backend)
    trace["rounds"] = [dict(r, score=r["score"]) for r in rounds]  # the rounds are this side's own; the graph is the pin
    check_against_trace(trace, (got_graph, got_best), rounds)
    assert [m.string for m in got_best] == trace["best_candidates"]
