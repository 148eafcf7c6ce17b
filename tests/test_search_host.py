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
