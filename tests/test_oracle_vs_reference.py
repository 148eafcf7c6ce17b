"""Cross-check oracle/restate.py against the REAL reference on fresh random inputs.  Runs only where
/root/reference is mounted (this container); skipped on the GPU box."""
import random

import numpy as np
import pytest

from oracle import ref_shim, restate as O

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def nm():
    return ref_shim.load_reference()


def _seq(rng, n, n_rate=0.0):
    s = rng.choice(list("ACGT"), size=n)
    if n_rate:
        s[rng.random(n) < n_rate] = "N"
    return "".join(s)


def test_scan_and_join(nm):
    rng = np.random.default_rng(99)
    seq = _seq(rng, 40000, 0.003)
    arr = np.frombuffer(seq.encode(), dtype=np.uint8)
    for motif, mp in (("GATC", 1), ("A", 0), ("CC[AT]GG", 1), ("G[AG].GAAG[CT]", 5), ("GCAC......GTT", 2), ("T.[ACG]", 2)):
        want = nm.utils.subseq_indices(motif, seq)
        np.testing.assert_array_equal(O.subseq_indices(motif, seq), want)
        np.testing.assert_array_equal(O.subseq_indices_np(motif, arr), want)
        sites = rng.choice(len(seq), size=4000, replace=False)
        meth, non = sites[:1500], sites[1500:]
        r = nm.find_motifs_bin.methylated_motif_occourances(nm.motif.Motif(motif, mp), seq, meth, non)
        g = O.methylated_motif_occourances(motif, mp, arr, meth, non, fast=True)
        np.testing.assert_array_equal(g[0], r[0])
        np.testing.assert_array_equal(g[1], r[1])


def test_scores(nm):
    rng = np.random.default_rng(5)
    for _ in range(200):
        a, b = int(rng.integers(0, 5000)), int(rng.integers(0, 5000))
        c, d = a + int(rng.integers(0, 5000)), b + int(rng.integers(0, 100000))
        nxt, cur = nm.model.BetaBernoulliModel(), nm.model.BetaBernoulliModel()
        nxt.update(a, b)
        cur.update(c, d)
        want = nm.find_motifs_bin.predictive_evaluation_score(nxt, cur)
        assert O.predictive_evaluation_score(O.posterior(a, b), O.posterior(c, d)) == pytest.approx(want, rel=1e-13, abs=1e-15)


def test_growth_functions(nm):
    rng = np.random.default_rng(3)
    seq = _seq(rng, 8000, 0.004)
    D = nm.seq.DNAsequence(seq)
    idx = sorted(rng.choice(len(seq), size=400, replace=False).tolist())
    ref_w = [s.sequence for s in D.sample_at_indices(idx, 20).sequences]
    assert O.sample_at_indices(seq, idx, 20) == ref_w
    ref_rc = [s.sequence for s in D.sample_at_indices(idx, 20).reverse_compliment().sequences]
    assert [O.reverse_complement_seq(w) for w in ref_w] == ref_rc
    es = nm.seq.EqualLengthDNASet([nm.seq.DNAsequence(w) for w in ref_w])
    arr = es.convert_to_DNAarray()
    np.testing.assert_array_equal(O.one_hot_windows(ref_w), np.asarray(arr))
    np.testing.assert_array_equal(O.background_pssm(ref_w), es.pssm())
    mask = nm.motif.Motif("." * 19 + "[AG]A" + "." * 20, 20).one_hot()
    for keep in (True, False):
        want = arr.filter_sequence_matches(mask, keep_matches=keep)
        _, got = O.filter_sequence_matches(np.asarray(arr), mask, keep)
        np.testing.assert_array_equal(got, np.asarray(want))
    random.seed(4)
    want_bg = [s.sequence for s in D.sample_n_subsequences_unique(41, 60, "C").sequences]
    assert O.sample_background(seq, 41, 60, "C", random.Random(4)) == want_bg


def test_motif_type_and_reverse_compliment(nm):
    """utils.motif_type / seq.reverse_compliment of the real reference on random IUPAC motifs."""
    from nanomotif_b200 import tables

    rng = np.random.default_rng(5)
    letters = list("ACGTRYSWKMBDHVN")
    for _ in range(300):
        L = int(rng.integers(1, 16))
        m = "".join(rng.choice(letters, size=L, p=[0.15] * 4 + [0.02] * 10 + [0.2]))
        assert O.motif_type(m) == nm.utils.motif_type(m) == tables.motif_type(m), m
        assert tables.reverse_compliment(m) == nm.seq.reverse_compliment(m)
