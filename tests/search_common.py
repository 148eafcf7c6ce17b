"""Shared helpers of the search tests: rebuild the golden search input from its seed and compare a search
result with tests/golden/search_trace.json (recorded from the REAL reference search loop)."""
import json
import os
import random

import numpy as np

from oracle import restate as O

HERE = os.path.dirname(os.path.abspath(__file__))


def load_trace(name="search_trace.json"):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def build_inputs(spec):
    """The golden search input, regenerated from its seed; spec["fasta"] names a FASTA fixture under tests/golden
    (cfg 1: the reference's bundled geobacillus plasmids) instead of random contigs."""
    from nanomotif_b200 import synth

    rng = np.random.default_rng(spec["seed"])
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    if "fasta" in spec:
        from nanomotif_b200.dataload import load_fasta

        named = [(k, np.frombuffer(v.encode(), dtype=np.uint8)) for k, v in
                 load_fasta(os.path.join(HERE, "golden", spec["fasta"])).items()]
    else:
        # a generator: sequence and pileup draws of one contig interleave on the same rng stream
        named = ((f"contig_{i}", synth.random_sequence(rng, L, spec["gc"], 2e-5)) for i, L in enumerate(spec["contig_lengths"]))
    for name, seq in named:
        contigs[name] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=spec["depth"], mod_types=(spec["mod_type"],),
                               planted=[tuple(x) for x in spec["planted"]])
        cols["contig"].append(np.full(len(p["position"]), name, dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(np.where(p["strand"] == 0, "+", "-").astype(object))
        cols["fraction_mod"].append(p["fraction_mod"])
    return contigs, {k: np.concatenate(v) for k, v in cols.items()}


class OracleBackend:
    """CPU backend for the search coroutines (TEST infrastructure: exercises the host-side search logic
    without a GPU).  Windows / PSSM / counts come from oracle/restate.py."""

    def __init__(self, contigs, pile, spec):
        from nanomotif_b200.model import BetaBernoulliModel

        self.Model = BetaBernoulliModel
        self.contigs, self.pile, self.spec = contigs, pile, spec
        pad, high = spec["padding"], spec["high"]
        windows = []
        for name, seq in contigs.items():
            sel = (pile["contig"] == name) & (pile["fraction_mod"] >= high)
            plus = pile["position"][sel & (pile["strand"] == "+")].tolist()
            minus = pile["position"][sel & (pile["strand"] == "-")].tolist()
            windows += O.methylation_windows(seq, plus, minus, pad)
        self.arr = O.one_hot_windows(windows)
        self.remaining = self.arr
        rng = random.Random(spec["random_seed"])
        bg = []
        for name, seq in contigs.items():
            bg += O.sample_background(seq, 2 * pad + 1, O.n_background_samples(len(seq)),
                                      O.MOD_TYPE_TO_CANONICAL[spec["mod_type"]], rng)
        self.bin_pssm = O.background_pssm(bg)
        self.calls = 0

    def handle(self, request):
        kind, arg = request
        if kind == "score":
            out = []
            for m in arg:
                self.calls += 1
                a, b = O.motif_model_bin(self.pile["contig"], self.pile["position"], self.pile["strand"],
                                         self.pile["fraction_mod"], self.contigs, m.string, m.mod_position,
                                         self.spec["low"], self.spec["high"], fast=True)
                mdl = self.Model()
                mdl.update(a, b)
                out.append(mdl)
            return out
        if kind == "expand":
            _, active = O.filter_sequence_matches(self.remaining, O.motif_one_hot(arg.string), True)
            return None if active is None else (active.shape[0], O.pssm(active))
        _, self.remaining = O.filter_sequence_matches(self.remaining, O.motif_one_hot(arg.string), False)
        return None if self.remaining is None else self.remaining.shape[0]


def check_against_trace(trace, result, rounds):
    graph, best = result
    assert [m.string for m in best[:len(trace["best_candidates"])]] == trace["best_candidates"]
    assert sorted(m.string for m in best[len(trace["best_candidates"]):]) == trace["missed"]
    want_rounds = [r for r in trace["rounds"] if "score" in r]
    assert len(rounds) == len(want_rounds)
    for got, want in zip(rounds, want_rounds):
        assert got["naive"] == want["naive"] and got["remaining"] == want["remaining"]
        assert got.get("kept") == want.get("kept")
        assert abs(got["score"] - want["score"]) <= 1e-6 * max(1.0, abs(want["score"]))
    nodes = {m.string: d for m, d in graph.nodes.items()}
    assert set(nodes) == {n["motif"] for n in trace["nodes"]}
    for n in trace["nodes"]:
        d = nodes[n["motif"]]
        assert (d["model"]._alpha, d["model"]._beta) == (n["alpha"], n["beta"]), n["motif"]
        assert abs(d["score"] - n["score"]) <= 1e-6 * max(1.0, abs(n["score"])), n["motif"]
        assert abs(d["priority"] - n["priority"]) <= 1e-9 * max(1.0, abs(n["priority"]))
        assert d["depth"] == n["depth"] and bool(d["visited"]) == n["visited"]
    assert sorted([u.string, v.string] for u, v in graph.edges()) == sorted(trace["edges"])


def load_merge_trace():
    with open(os.path.join(HERE, "golden", "merge_trace.json")) as f:
        return json.load(f)


def merge_inputs(golden, run, Motif, Model):
    """(rows, clusters) of one golden merge run as nanomotif_b200 objects."""
    rows = [dict(motif=s, mod_position=p, score=1.0, reference="bin1", mod_type="a", model=Model()) for s, p in golden["motifs"]]
    mk = lambda sp: Motif(sp[0], sp[1])
    clusters = [(mk(c["merged"]), [mk(m) for m in c["premerge"]], {mk(m) for m in c["pre_variants"]},
                 {mk(m) for m in c["new_variants"]}) for c in run["clusters"]]
    return rows, clusters


def check_merge_against_trace(run, rows_out, decisions):
    assert len(decisions) == len(run["decisions"])
    for got, want in zip(decisions, run["decisions"]):
        assert got["scored"] == want["scored"] and got["accepted"] == want["accepted"]
        if want["scored"]:
            assert [got["merge_model"]._alpha, got["merge_model"]._beta] == want["merge_model"]
            assert [got["variants_model"]._alpha, got["variants_model"]._beta] == want["variants_model"]
            assert abs(got["merge_score"] - want["merge_score"]) <= 1e-9 * max(1.0, abs(want["merge_score"]))
    assert [(r["motif"], r["mod_position"]) for r in rows_out] == [(r["motif"], r["mod_position"]) for r in run["rows"]]
    for got, want in zip(rows_out, run["rows"]):
        if want["merged"]:
            assert [got["model"]._alpha, got["model"]._beta] == want["model"]
            assert abs(got["score"] - want["score"]) <= 1e-9 * max(1.0, abs(want["score"]))
            assert got["reference"] == "bin1" and got["mod_type"] == "a"
