"""Host-side search logic (nanomotif_b200/search.py) against the trace of the REAL reference search loop,
driven by a CPU oracle backend (no GPU)."""
import numpy as np

from nanomotif_b200 import search
from search_common import OracleBackend, build_inputs, check_against_trace, load_trace


def test_search_reproduces_reference_trace():
    trace = load_trace()
    spec = trace["spec"]
    contigs, pile = build_inputs(spec)
    backend = OracleBackend(contigs, pile, spec)
    assert backend.arr.shape[0] == trace["total_windows"]
    np.testing.assert_array_equal(backend.bin_pssm, np.array(trace["bin_pssm"]))
    rounds = []
    co = search.find_candidates(spec["mod_type"], spec["padding"], backend.bin_pssm, backend.arr.shape[0],
                                min_kl=spec["min_kl"], score_threshold=spec["score_threshold"], trace=rounds)
    result = search.run(co, backend)
    check_against_trace(trace, result, rounds)
    assert backend.calls == trace["scoring_calls"]  # the same motif_model_bin evaluations, batched differently


def test_cfg1_bundled_assembly_recovers_golden_motifs():
    """BASELINE.json configs[0]: the reference's bundled geobacillus plasmids (tests/golden/*.fasta) with the shipped
    golden motifs planted (nanomotif/datasets/geobacillus-plasmids.bin-motifs.tsv:2-5).  The trace was recorded from
    the REAL MotifSearcher (tests/golden/generate_search_golden.py --cfg1)."""
    trace = load_trace("search_trace_cfg1.json")
    spec = trace["spec"]
    contigs, pile = build_inputs(spec)
    assert {k: len(v) for k, v in contigs.items()} == {"contig_3": 82915, "contig_2": 93311}
    backend = OracleBackend(contigs, pile, spec)
    assert backend.arr.shape[0] == trace["total_windows"]
    rounds = []
    co = search.find_candidates(spec["mod_type"], spec["padding"], backend.bin_pssm, backend.arr.shape[0],
                                min_kl=spec["min_kl"], score_threshold=spec["score_threshold"], trace=rounds)
    graph, best = search.run(co, backend)
    check_against_trace(trace, (graph, best), rounds)
    found = [m.new_stripped_motif() for m in best]
    for want in (search.Motif("GATC", 1), search.Motif("ACCCA", 4), search.Motif("CCAAAT", 4)):
        assert want in found
    grn = search.Motif("G[AG].GAAG[CT]", 5)  # GRNGAAGY comes out as its concrete variants (merged later, motif.py:470-560)
    rest = [m for m in found if m.string not in ("GATC", "ACCCA", "CCAAAT")]
    assert rest and all(m.mod_position == 5 and m.length() == grn.length() for m in rest)
    for m in rest:  # every constrained position agrees with GRNGAAGY
        assert all(t == "." or set(t.strip("[]")) <= set(g.strip("[]")) for t, g in zip(m.split(), grn.split())), m


def test_lockstep_driver_matches_sequential():
    trace = load_trace()
    spec = trace["spec"]
    contigs, pile = build_inputs(spec)
    searches, traces = [], []
    for _ in range(2):
        b = OracleBackend(contigs, pile, spec)
        t = []
        traces.append(t)
        searches.append((search.find_candidates(spec["mod_type"], spec["padding"], b.bin_pssm, b.arr.shape[0],
                                                min_kl=spec["min_kl"], score_threshold=spec["score_threshold"], trace=t), b))
    batches = []

    def batch_score(reqs):
        batches.append(len(reqs))
        return [b.handle(("score", motifs)) for b, motifs in reqs]

    results = search.run_lockstep(searches, batch_score)
    for res, t in zip(results, traces):
        check_against_trace(trace, res, t)
    assert max(batches) == 2


def test_motif_graph_queries():
    g = search.MotifGraph()
    M = search.Motif
    a, b, c, d = M("A", 0), M("GA", 1), M("GAT", 1), M("CA", 1)
    for n, s in ((a, 0.0), (b, 5.0), (c, 6.0), (d, 4.0)):
        g.add_node(n, score=s)
    g.add_edge(a, b)
    g.add_edge(b, c)
    g.add_edge(a, d)
    assert g.ancestors(c) == {a, b} and g.descendants(a) == {b, c, d}
    assert g.get_missed_candidates([c], 3) == {d}          # b has a kept descendant, c is kept
    assert g.get_missed_candidates([], 3) == {b, d}        # c has a high-scoring ancestor


def test_merge_group_reproduces_reference_decisions():
    """search.merge_group (two batched score requests per group) against the golden trace of the reference's
    merge_motifs_in_df body (real merge_motifs / predictive_evaluation_score / get_parent_scores)."""
    from nanomotif_b200.model import BetaBernoulliModel
    from search_common import check_merge_against_trace, load_merge_trace, merge_inputs

    golden = load_merge_trace()
    spec = load_trace()["spec"]
    contigs, pile = build_inputs(spec)
    for run in golden["runs"]:
        backend = OracleBackend(contigs, pile, spec)
        rows, clusters = merge_inputs(golden, run, search.Motif, BetaBernoulliModel)
        requests = []

        class Counting:
            def handle(self, request):
                requests.append(len(request[1]))
                return backend.handle(request)

        decisions = []
        out = search.run(search.merge_group(rows, clusters, run["merge_threshold"], trace=decisions), Counting())
        check_merge_against_trace(run, out, decisions)
        n_accepted = sum(d["accepted"] for d in run["decisions"])
        # the reference scores an accepted motif twice (:1502 and inside get_parent_scores :1400); one request holds it once
        assert len(requests) == 2 and sum(requests) == run["scoring_calls"] - n_accepted
    # several groups in lock-step: the score requests of all groups arrive together
    run = golden["runs"][0]
    groups, batches = [], []
    for _ in range(3):
        rows, clusters = merge_inputs(golden, run, search.Motif, BetaBernoulliModel)
        groups.append((rows, clusters, OracleBackend(contigs, pile, spec), run["merge_threshold"]))

    def batch_score(reqs):
        batches.append(len(reqs))
        return [b.handle(("score", motifs)) for b, motifs in reqs]

    for out in search.merge_motifs_in_groups(groups, batch_score):
        check_merge_against_trace(run, out, run["decisions"] and [dict(d, merge_model=_M(d.get("merge_model")), variants_model=_M(d.get("variants_model"))) for d in run["decisions"]])
    assert batches == [3, 3]


class _M:
    def __init__(self, ab):
        self._alpha, self._beta = ab if ab else (None, None)
