#!/usr/bin/env python
"""Generate tests/golden/search_trace.json by running the REAL reference search loop.

    python tests/golden/generate_search_golden.py          (needs /root/reference)
    python tests/golden/generate_search_golden.py --cfg1   -> search_trace_cfg1.json: BASELINE.json configs[0], the
        bundled geobacillus plasmids (nanomotif/datasets/geobacillus-plasmids.assembly.fasta, copied next to this
        script as a data fixture) with a synthetic depth-100 pileup planting the shipped golden motifs
        (nanomotif/datasets/geobacillus-plasmids.bin-motifs.tsv:2-5) -- the real pileup is not in the tree

What is real and what is restated:
  * REAL (imported unchanged through oracle/ref_shim.py): MotifSearcher.run (best-first expansion,
    find_motifs_bin.py:1026-1182), _motif_child_nodes_kl_dist_max, _priority_function,
    predictive_evaluation_score, get_parent_scores (:1382-1433), Motif, MotifTree (incl.
    get_missed_candidates), DNAsequence / EqualLengthDNASet / DNAarray (windows, filter, pssm),
    BetaBernoulliModel.
  * PATCHED: nanomotif.find_motifs_bin.motif_model_bin -- the reference version filters a polars frame
    (not installed here); it is replaced by oracle.restate.motif_model_bin on the same columns
    (pinned separately against the reference's scan/join functions).
  * RESTATED in this script: the head of find_best_candidates (polars filters -> numpy, :625-686) and the
    glue of its candidate loop (:695-834), calling the real functions above.
The input is regenerated from a seed by nanomotif_b200.synth (numpy PCG64), so only the seed travels.
"""
import json
import math
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import restate as O  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

nm = load_reference()
fmb = nm.find_motifs_bin
Motif = nm.motif.Motif

SPEC = dict(seed=2026, contig_lengths=[180000, 90000], gc=0.5, depth=20, mod_type="a",
            planted=[["GATC", 1, "a"], ["GCAC......GTT", 2, "a"], ["CAA..[AT]TG", 2, "a"]], padding=20, low=0.3, high=0.7,
            min_kl=0.05, score_threshold=1.5, random_seed=1)


CFG1_SPEC = dict(seed=1, fasta="geobacillus-plasmids.assembly.fasta", depth=100, mod_type="a",
                 planted=[["ACCCA", 4, "a"], ["CCAAAT", 4, "a"], ["G[AG].GAAG[CT]", 5, "a"], ["GATC", 1, "a"]], padding=20,
                 low=0.3, high=0.7, min_kl=0.05, score_threshold=1.5, random_seed=1)


def read_fasta(path):
    out, name = {}, None
    for line in open(path):
        if line.startswith(">"):
            name = line[1:].split()[0]
            out[name] = []
        else:
            out[name].append(line.strip().upper())
    return {k: "".join(v) for k, v in out.items()}


def build_inputs(spec):
    from nanomotif_b200 import synth

    rng = np.random.default_rng(spec["seed"])
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    if "fasta" in spec:
        named = [(k, np.frombuffer(v.encode(), dtype=np.uint8)) for k, v in
                 read_fasta(os.path.join(os.path.dirname(os.path.abspath(__file__)), spec["fasta"])).items()]
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


class Cols(dict):
    def is_empty(self):
        return len(self["position"]) == 0


def main():
    cfg1 = "--cfg1" in sys.argv
    spec = CFG1_SPEC if cfg1 else SPEC
    contigs, pile = build_inputs(spec)
    pile = Cols(pile)
    pad, low, high = spec["padding"], spec["low"], spec["high"]
    calls = {"n": 0}

    def motif_model_bin(pileup, contigs, motif, model, low_meth_threshold, high_meth_threshold):
        calls["n"] += 1
        a, b = O.motif_model_bin(pileup["contig"], pileup["position"], pileup["strand"], pileup["fraction_mod"],
                                 {k: v.sequence for k, v in contigs.items()}, motif.string, motif.mod_position,
                                 low_meth_threshold, high_meth_threshold, fast=True)
        model.update(a, b)
        return model

    fmb.motif_model_bin = motif_model_bin
    bin_sequences = {k: nm.seq.DNAsequence(v) for k, v in contigs.items()}

    # ---- head of find_best_candidates (:625-686); contig order = dict order ----
    random.seed(spec["random_seed"])
    conf = pile["fraction_mod"] >= high
    meth, background = None, None
    for name, dseq in bin_sequences.items():
        sel = conf & (pile["contig"] == name)
        n_samples = int(max(math.ceil(len(dseq) * 0.01), 50))
        index_plus = pile["position"][sel & (pile["strand"] == "+")].tolist()
        index_minus = pile["position"][sel & (pile["strand"] == "-")].tolist()
        sample = dseq.sample_n_subsequences_unique(pad * 2 + 1, n_samples, base=nm.constants.MOD_TYPE_TO_CANONICAL[spec["mod_type"]])
        background = sample if background is None else background + sample.sequences
        strings = []
        t = dseq.sample_at_indices(index_plus, pad)
        if t is not None:
            strings += t.sequences
        t = dseq.sample_at_indices(index_minus, pad)
        if t is not None:
            strings += t.reverse_compliment().sequences
        es = nm.seq.EqualLengthDNASet(strings)
        meth = es if meth is None else meth + es
    methylation_sequences = meth.convert_to_DNAarray()
    total_seqs = methylation_sequences.shape[0]
    bin_pssm = background.pssm()

    # ---- candidate loop (:688-834) ----
    root = Motif("." * pad + nm.constants.MOD_TYPE_TO_CANONICAL[spec["mod_type"]] + "." * pad, pad)
    clone = methylation_sequences.copy()
    best_candidates, dead_ends, graph, rounds = [], 0, None, []
    while True:
        if dead_ends >= 25:
            break
        searcher = fmb.MotifSearcher(root, bin_sequences, bin_pssm, pile, clone, pad, high, low, motif_graph=graph,
                                     min_kl=spec["min_kl"], max_rounds_since_new_best=30)
        graph, naive = searcher.run()
        if naive == root:
            rounds.append(dict(naive=naive.string, stop="root"))
            break
        temp, prune, single = naive, set(), False
        while True:
            parents = fmb.get_parent_scores(temp, pile, bin_sequences, low, high)
            scores = [d["score"] for d in parents.values()]
            mean_score = np.mean(scores)
            for parent, d in parents.items():
                if d["score"] < 0.4:
                    prune.add(d["motif_position"])
            if len(prune) == 0:
                break
            ms = temp.split()
            for i in prune:
                ms[i] = "."
            pruned = Motif("".join(ms), temp.mod_position)
            if len(pruned.string.replace(".", "")) == 1:
                single = True
                break
            if pruned == temp:
                break
            temp = pruned
        if single or mean_score < spec["score_threshold"]:
            graph.nodes[naive]["score"] = np.mean([d["score"] for d in parents.values()])
        elif temp != naive:
            child_model = [d["child_model"] for _, d in parents.items()][0]
            graph.add_node(temp, model=child_model, motif=temp, visited=True, score=mean_score, priority=0, depth=0)
            naive = temp
        else:
            graph.nodes[naive]["score"] = np.mean([d["score"] for d in parents.values()])
        before = clone.shape[0]
        clone = clone.filter_sequence_matches(naive.one_hot(), keep_matches=False)
        rec = dict(naive=naive.string, score=float(graph.nodes[naive]["score"]), before=int(before),
                   remaining=None if clone is None else int(clone.shape[0]))
        if clone is None:
            rec["stop"] = "no sequences"
            rounds.append(rec)
            break
        if graph.nodes[naive]["score"] < spec["score_threshold"]:
            dead_ends += 1
            rec["kept"] = False
            rounds.append(rec)
            continue
        rec["kept"] = True
        rounds.append(rec)
        best_candidates.append(naive)
        if clone.shape[0] / total_seqs < 0.001:
            break
    missed = graph.get_missed_candidates(best_candidates, spec["score_threshold"])
    missed = [c for c in missed if not c.sub_motif_of_any(best_candidates) or not any(b.sub_motif_of(c) for b in best_candidates)]
    out = dict(spec=spec, total_windows=int(total_seqs), bin_pssm=bin_pssm.tolist(), rounds=rounds,
               best_candidates=[c.string for c in best_candidates], missed=sorted(c.string for c in missed),
               scoring_calls=calls["n"],
               nodes=[dict(motif=n.string, alpha=int(d["model"]._alpha), beta=int(d["model"]._beta), score=float(d["score"]),
                           priority=float(d["priority"]), depth=int(d["depth"]), visited=bool(d["visited"]))
                      for n, d in graph.nodes(data=True)],
               edges=[[u.string, v.string] for u, v in graph.edges()])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "search_trace_cfg1.json" if cfg1 else "search_trace.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["nodes"]), "nodes,", calls["n"], "scoring calls")
    print("best:", [c.strip(".") for c in out["best_candidates"]], "missed:", [c.strip(".") for c in out["missed"]])
    for r in rounds:
        print({k: (v.strip(".") if isinstance(v, str) else v) for k, v in r.items()})


if __name__ == "__main__":
    main()
