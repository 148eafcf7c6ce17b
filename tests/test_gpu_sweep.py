"""cfg 5 (exhaustive candidate sweep) as a parity case: a seeded sample of IUPAC 4-8-mers with the canonical
base at an admissible modified position, plus bipartite X{3,4} N{4..8} Y{3,4} motifs, against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import restate as O

IUPAC = "ACGTRYSWKMBDHVN"


def _sample_motifs(rng, n):
    out = []
    while len(out) < n:
        if rng.random() < 0.75:
            k = int(rng.integers(4, 9))
            s = "".join(rng.choice(list(IUPAC), size=k))
            if s[0] == "N" or s[-1] == "N":  # canonical form: no flanking wildcard
                continue
        else:
            s = "".join(rng.choice(list("ACGT"), size=int(rng.integers(3, 5)))) + "N" * int(rng.integers(4, 9)) + \
                "".join(rng.choice(list("ACGT"), size=int(rng.integers(3, 5))))
        pos = [i for i, ch in enumerate(s) if ch == "A"]
        if not pos:
            continue
        out.append((s, int(rng.choice(pos))))
    return out


def test_iupac_sweep_sample_matches_oracle():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nanomotif_b200 as nmb
    from nanomotif_b200 import synth

    rng = np.random.default_rng(15)
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    for i, L in enumerate((120000, 60000, 2600)):
        seq = synth.random_sequence(rng, L, 0.4 + 0.1 * i, 5e-5)
        contigs[f"c{i}"] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=12, mod_types=("a",))
        cols["contig"].append(np.full(len(p["position"]), f"c{i}", dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(np.where(p["strand"] == 0, "+", "-"))
        cols["fraction_mod"].append(p["fraction_mod"])
    pile = {k: np.concatenate(v) for k, v in cols.items()}
    sample = _sample_motifs(rng, 300)
    motifs = [nmb.Motif(s, p).from_iupac() for s, p in sample]  # IUPAC letters with set semantics -> regex classes
    scorer = nmb.BinScorer(pile, contigs, 0.3, 0.7)
    got = scorer.score(motifs)
    assert got.shape == (300, 2)
    nonzero = 0
    for (s, p), m, g in zip(sample, motifs, got):
        want = O.motif_model_bin(pile["contig"], pile["position"], pile["strand"], pile["fraction_mod"], contigs,
                                 m.string, m.mod_position, fast=True)
        assert tuple(g) == want, (s, p)
        nonzero += sum(want) > 0
    assert nonzero > 100


def test_exhaustive_tables_equal_per_motif_counts():
    """K8: the histogram + subset-sum tables hold, for EVERY IUPAC motif of length 4..8, exactly the counts that a scan of
    that motif gives -- checked on a seeded sample against the CPU oracle (small assembly with non-ACGT letters, contigs
    shorter than a window, contig ends) and against K2 for whole tables' worth of motifs."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nanomotif_b200 as nmb
    from nanomotif_b200 import synth
    from nanomotif_b200.sweep import SweepIndex

    rng = np.random.default_rng(16)
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    for i, L in enumerate((90000, 40000, 2600, 7, 5, 70000, 3)):
        seq = synth.random_sequence(rng, L, 0.4 + 0.03 * i, 2e-4 if L > 100 else 0.0)
        if L > 1000:
            seq[rng.integers(0, L, 20)] = ord("N")  # isolated non-ACGT letters: the wildcard matches them, sets do not
            seq[-1] = ord("R")
        contigs[f"c{i}"] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=12, mod_types=("a",))
        cols["contig"].append(np.full(len(p["position"]), f"c{i}", dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(np.where(p["strand"] == 0, "+", "-"))
        cols["fraction_mod"].append(p["fraction_mod"])
    pile = {k: np.concatenate(v) for k, v in cols.items()}
    scorer = nmb.BinScorer(pile, contigs, 0.3, 0.7)
    index = SweepIndex(scorer.assembly, scorer.pileup, 0).add()
    sample = [(s, p) for s, p in _sample_motifs(rng, 400) if 4 <= len(s) <= 8]
    sample += [("ANNNA", 0), ("NANN", 1), ("NNNNNNNA", 7), ("ANNNNNNN", 0), ("GATC", 1), ("TTAA", 3), ("NNANN", 2)]
    assert len(sample) > 180
    nonzero = 0
    for s, p in sample:
        got = index.counts(s, p)
        ref = _oracle_counts(pile, contigs, s, p)
        assert got == ref, (s, p, got, ref)
        nonzero += sum(ref) > 0
    assert nonzero > 100
    # one whole table against K2: all 15^3 motifs of length 4 with the modified A at position 1
    n_mod, n_nomod = index.table(4, 1, "A")
    all4 = [SweepIndex.index_motif(i, 4, 1, "A") for i in range(15 ** 3)]
    inner = [(i, s) for i, s in enumerate(all4) if s[0] != "N" and s[-1] != "N"]  # K2 strips flanking wildcards
    got = scorer.score([nmb.Motif(s, 1).from_iupac() for _, s in inner])
    sel = torch.tensor([i for i, _ in inner], device=n_mod.device)
    np.testing.assert_array_equal(got[:, 0], n_mod[sel].cpu().numpy())
    np.testing.assert_array_equal(got[:, 1], n_nomod[sel].cpu().numpy())
    # whole tables against the oracle's table computation (histogram + subset sums in numpy): small contigs only
    small = {n: s for n, s in contigs.items() if len(s) < 3000}
    sel_rows = np.isin(pile["contig"], list(small))
    sub = {k: v[sel_rows] for k, v in pile.items()}
    small_scorer = nmb.BinScorer(sub, small, 0.3, 0.7)
    small_index = SweepIndex(small_scorer.assembly, small_scorer.pileup, 0).add()
    for k, mp in ((4, 0), (5, 2), (6, 5)):
        want_mod, want_nomod = O.sweep_table(small, sub["contig"], sub["position"], sub["strand"], sub["fraction_mod"], k, mp)
        got_mod, got_nomod = small_index.table(k, mp, "A")
        np.testing.assert_array_equal(got_mod.cpu().numpy(), want_mod)
        np.testing.assert_array_equal(got_nomod.cpu().numpy(), want_nomod)
        assert want_mod.sum() + want_nomod.sum() > 0
    # bipartite shapes X{3,4} N{4..8} Y{3,4}: a sample against the oracle, one whole (3, 6, 3) table against K2
    index.add_bipartite()
    hits = 0
    for _ in range(120):
        a, g, b = int(rng.integers(3, 5)), int(rng.integers(4, 9)), int(rng.integers(3, 5))
        left, right = "".join(rng.choice(list("ACGT"), a)), "".join(rng.choice(list("ACGT"), b))
        s = left + "N" * g + right
        pos = [i for i, ch in enumerate(s) if ch == "A"]
        if not pos:
            continue
        p = int(rng.choice(pos))
        got, ref = index.bipartite_counts(left, g, right, p), _oracle_counts(pile, contigs, s, p)
        assert got == ref, (s, p, got, ref)
        hits += sum(ref) > 0
    assert hits > 30
    bip_mod, bip_nomod = index.bipartite_table(3, 6, 3, 1)
    motifs, idx = [], []
    for l0 in "ACGT":
        for l2 in "ACGT":
            for r in ("".join(x) for x in __import__("itertools").product("ACGT", repeat=3)):
                left = l0 + "A" + l2
                motifs.append(nmb.Motif(left + "N" * 6 + r, 1).from_iupac())
                idx.append(SweepIndex.bipartite_index(left, r))
    got = scorer.score(motifs)
    sel = torch.tensor(idx, device=bip_mod.device)
    np.testing.assert_array_equal(got[:, 0], bip_mod[sel].cpu().numpy())
    np.testing.assert_array_equal(got[:, 1], bip_nomod[sel].cpu().numpy())
    assert int(got.sum()) > 1000
    # whole tables of the other shapes (they come from marginalising (4, g, 4) + edge windows) against K2: every motif
    # with an A at the modified position, in the left part and in the right part
    import itertools

    for a, g, b, mp in ((4, 4, 3, 2), (3, 8, 4, 3 + 8 + 1), (4, 5, 4, 0), (3, 4, 3, 3 + 4 + 2)):
        t_mod, t_nomod = index.bipartite_table(a, g, b, mp)
        motifs, idx = [], []
        for letters in itertools.product("ACGT", repeat=a + b - 1):
            concrete = list(letters)
            k = mp if mp < a else mp - g
            concrete.insert(k, "A")
            left, right = "".join(concrete[:a]), "".join(concrete[a:])
            motifs.append(nmb.Motif(left + "N" * g + right, mp).from_iupac())
            idx.append(SweepIndex.bipartite_index(left, right))
        got = scorer.score(motifs)
        sel = torch.tensor(idx, device=t_mod.device)
        np.testing.assert_array_equal(got[:, 0], t_mod[sel].cpu().numpy(), err_msg=str((a, g, b, mp)))
        np.testing.assert_array_equal(got[:, 1], t_nomod[sel].cpu().numpy(), err_msg=str((a, g, b, mp)))
        assert int(got.sum()) > 100
    # candidates: the planted GATC stands out among all 4-mers
    cand = index.candidates(4, 1, "A", min_mean=0.8, min_mod=100)
    assert ("GATC", int(n_mod[SweepIndex.motif_index("GATC", 1)]), int(n_nomod[SweepIndex.motif_index("GATC", 1)])) in cand
    assert all(s[1] == "A" and s[0] != "N" and s[-1] != "N" for s, _, _ in cand)


def _oracle_counts(pile, contigs, iupac, mod_pos, low=0.3, high=0.7):
    """Counts of an IUPAC motif AS WRITTEN (a flanking N is a regex '.', which needs a character of the contig) with the
    reference's scan + join restated in oracle/restate.py; for motifs without flanking N this is motif_model_bin."""
    rx = "".join(O.IUPAC_TO_REGEX[c] for c in iupac)
    rc_rx, rc_pos = O.reverse_complement_motif(rx, mod_pos)
    n_mod = n_nomod = 0
    for name, seq in contigs.items():
        sel = pile["contig"] == name
        pos, strand, frac = pile["position"][sel], pile["strand"][sel], pile["fraction_mod"][sel]
        arr = np.frombuffer(seq.encode(), dtype=np.uint8)
        for st, motif, mp in (("+", rx, mod_pos), ("-", rc_rx, rc_pos)):
            on = strand == st
            a, b = O.methylated_motif_occourances(motif, mp, arr, pos[on & (frac >= high)], pos[on & (frac <= low)], True)
            n_mod += len(a)
            n_nomod += len(b)
    return n_mod, n_nomod
