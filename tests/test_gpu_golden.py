"""The CUDA path against the committed golden vectors of the real reference (tests/golden)."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
        return json.load(f)


def test_subseq_indices_golden(G):
    import nanomotif_b200 as nmb

    seen_literal = False
    for c in G["subseq_indices"]:  # includes the literal 'N' motif: regex-literal semantics (N matches N only)
        seen_literal |= "N" in c["motif"]
        assert nmb.subseq_indices(c["motif"], c["seq"]).tolist() == c["result"], c["motif"]
    assert seen_literal


def test_literal_letters_in_motifs_match_like_the_regex(G):
    """A non-ACGT letter typed into a motif is a regex literal (utils.py:61-66): it matches the same contig letter and
    nothing else; '.' matches everything.  Against the oracle's regex on contigs with N / R / Y runs."""
    import nanomotif_b200 as nmb
    from oracle import restate as O

    rng = np.random.default_rng(5)
    letters = np.frombuffer(b"ACGTNRY", dtype=np.uint8)
    for L in (50, 3000, 70000):
        seq = letters[rng.choice(7, size=L, p=[0.22, 0.22, 0.22, 0.22, 0.06, 0.03, 0.03])].tobytes().decode()
        for motif in ("N", "NN", "AN", "N.A", "A.N.C", "GNNT", "R", "[AC]N", ".N", "N.", "TNA[GT].N"):
            np.testing.assert_array_equal(nmb.subseq_indices(motif, seq), O.subseq_indices(motif, seq), err_msg=motif)
        meth = np.sort(rng.choice(L, size=L // 3, replace=False)).astype(np.int64)
        non = np.setdiff1d(np.arange(L), meth)[: L // 3].astype(np.int64)
        for motif, mp in (("AN", 0), ("N.A", 2), ("GNNT", 3), ("N", 0), (".NA.", 2)):
            got = nmb.methylated_motif_occourances(nmb.Motif(motif, mp), seq, meth, non)
            want = O.methylated_motif_occourances(motif, mp, seq, meth, non)
            np.testing.assert_array_equal(got[0], want[0], err_msg=motif)
            np.testing.assert_array_equal(got[1], want[1], err_msg=motif)


def test_methylated_motif_occourances_golden(G):
    import nanomotif_b200 as nmb

    for c in G["methylated_motif_occourances"]:
        seq = G["seq_a"] if c["seq"] == "seq_a" else c["seq"]
        a, b = nmb.methylated_motif_occourances(nmb.Motif(c["motif"], c["mod_pos"]), seq,
                                                np.array(c["meth"], dtype=np.int64), np.array(c["nonmeth"], dtype=np.int64))
        assert [a.tolist(), b.tolist()] == c["result"]  # unsorted input order is preserved


def test_growth_golden(G):
    import nanomotif_b200 as nmb
    from nanomotif_b200 import growth
    from nanomotif_b200.device import DeviceAssembly

    g = G["growth"]
    pad = g["padding"]
    asm = DeviceAssembly.from_sequences({"c": g["seq"]})
    pos = np.array(g["plus"] + g["minus"], dtype=np.int64)
    strand = np.array([0] * len(g["plus"]) + [1] * len(g["minus"]), dtype=np.uint8)
    arr = growth.methylation_windows(asm, np.zeros(len(pos), np.int32), pos, strand, np.ones(len(pos)), 0.7, pad)
    assert arr.shape == (len(g["windows"]), 2 * pad + 1, 4)
    assert arr.column_counts().tolist() == g["one_hot_sum"]
    np.testing.assert_array_equal(arr.pssm(), np.array(g["pssm_all"]))
    np.testing.assert_array_equal(arr.exact_pssm(), np.array(g["exact_pssm_all"]))
    random.seed(g["bg_seed"])
    starts = growth.sample_background_starts(g["seq"], 2 * pad + 1, g["bg_n"], g["bg_base"])
    assert [g["seq"][s:s + 2 * pad + 1] for s in starts] == g["bg_windows"]
    bg = growth.DeviceDNAarray.from_positions(asm, np.zeros(len(starts), np.int64), np.array(starts) + pad,
                                              np.zeros(len(starts), np.uint8), pad).exact_pssm()
    np.testing.assert_array_equal(bg, np.array(g["bin_pssm"]))
    for step in g["steps"]:
        m = nmb.Motif(step["motif"], step["mod_pos"])
        active = arr.copy().filter_sequence_matches(m.one_hot())
        assert active.shape[0] == step["n_active"]
        assert active.column_counts().tolist() == step["column_counts"]
        np.testing.assert_array_equal(active.pssm(), np.array(step["pssm"]))
        n_act, pssm, kl = arr.expand([m], bg)
        assert n_act[0] == step["n_active"]
        np.testing.assert_allclose(kl[0], np.array(step["kl"]), rtol=1e-6, atol=1e-12)
        children = growth.kl_children(m, active.pssm(), bg, kl=kl[0])
        assert [[c.string, c.mod_position] for c in children] == step["children"]
        rest = arr.filter_sequence_matches(m.one_hot(), keep_matches=False)
        assert (0 if rest is None else rest.shape[0]) == step["n_removed_rest"]


def test_drop_in_bin_model_functions_equal_the_reference_functions():
    """motif_model_bin / motif_model_contig (with its four position lists) / get_parent_scores on the GPU against
    tests/golden/binmodel_vectors.json, which tests/golden/generate_binmodel_golden.py recorded by running the reference's
    OWN functions (find_motifs_bin.py:1265-1331, 1382-1433) on a frame.  Every mismatch is reported, not just the first."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nanomotif_b200 as nmb
    from test_binmodel_golden import build_inputs, digest

    with open(os.path.join(HERE, "golden", "binmodel_vectors.json")) as f:
        V = json.load(f)
    contigs, pile = build_inputs(V["spec"])
    assert len(pile["position"]) == V["n_rows"] and int(pile["position"].sum()) == V["checksum"]
    sub = {mt: {k: v[pile["mod_type"] == mt] for k, v in pile.items()} for mt in V["spec"]["mod_types"]}
    bad = []
    for rec in V["bin"]:
        m = nmb.motif_model_bin(sub[rec["mod_type"]], contigs, nmb.Motif(rec["motif"], rec["mod_pos"]),
                                nmb.BetaBernoulliModel(), rec["low"], rec["high"])
        if list(m.get_raw_counts()) != rec["counts"]:
            bad.append(("bin", rec, list(m.get_raw_counts())))
    per_contig = {}
    for rec in V["contig"]:
        key = (rec["mod_type"], rec["contig"])
        if key not in per_contig:
            sel = sub[rec["mod_type"]]["contig"] == rec["contig"]
            per_contig[key] = {k: v[sel] for k, v in sub[rec["mod_type"]].items()}
        m, pos = nmb.motif_model_contig(per_contig[key], contigs[rec["contig"]], nmb.BetaBernoulliModel(),
                                        nmb.Motif(rec["motif"], rec["mod_pos"]), save_motif_positions=True)
        if list(m.get_raw_counts()) != rec["counts"] or {k: digest(v) for k, v in pos.items()} != rec["positions"]:
            bad.append(("contig", rec["motif"], rec["contig"], list(m.get_raw_counts()), rec["counts"]))
    for rec in V["parents"]:
        res = nmb.get_parent_scores(nmb.Motif(rec["motif"], rec["mod_pos"]), sub[rec["mod_type"]], contigs, 0.3, 0.7)
        got = [(k.string, int(k.mod_position), int(v["motif_position"]), list(v["parent_model"].get_raw_counts()),
                list(v["child_model"].get_raw_counts())) for k, v in res.items()]
        want = [(p["parent"], p["mod_pos"], p["motif_position"], p["parent_counts"], p["child_counts"]) for p in rec["parents"]]
        if got != want:
            bad.append(("parents", rec["motif"], got, want))
        else:
            for v, p in zip(res.values(), rec["parents"]):
                if v["score"] != pytest.approx(p["score"], rel=1e-6, abs=1e-9):
                    bad.append(("score", rec["motif"], p["parent"], float(v["score"]), p["score"]))
    assert not bad, bad[:6]
