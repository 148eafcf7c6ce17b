"""Parity of the device pileup filters and the bedMethyl loader against the oracle / reference KATs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import restate as O


@pytest.fixture(scope="module")
def dl():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nanomotif_b200 import dataload

    return dataload


def _table(rng, n_contigs=6, length=30000):
    from nanomotif_b200 import synth
    from nanomotif_b200.pileup import PileupTable

    cols = {k: [] for k in ("contig", "position", "strand", "fraction_mod", "mod_type", "Nvalid_cov")}
    for i in range(n_contigs):
        L = int(length * (0.2 + rng.random()))
        seq = synth.random_sequence(rng, L, 0.5)
        planted = synth.DEFAULT_PLANTED if i % 3 else ()
        p = synth.synth_pileup(seq, rng, depth=int(rng.integers(4, 12)), planted=planted)
        n = len(p["position"])
        cols["contig"].append(np.full(n, f"contig_{i}", dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(np.where(p["strand"] == 0, "+", "-").astype(object))
        cols["fraction_mod"].append(p["fraction_mod"])
        cols["mod_type"].append(np.array(synth.MOD_TYPES, dtype=object)[p["mod_type"]])
        cols["Nvalid_cov"].append(p["Nvalid_cov"])
    c = {k: np.concatenate(v) for k, v in cols.items()}
    return PileupTable(c["contig"], c["position"], c["strand"], c["fraction_mod"], c["mod_type"], c["Nvalid_cov"])


def test_adjacency_reference_kat(dl):
    # /root/reference/tests/test_dataload.py:37-69
    frac = [0.8, 0.9, 0.1, 0.95, 0.85, 0.2, 0.75, 0.9, 0.05, 0.8]
    d = dict(contig=np.array(["contig1"] * 10), position=list(range(10)), mod_type=np.array(["m6A"] * 10),
             strand=np.array(["+"] * 10), fraction_mod=frac, Nvalid_cov=[10] * 10)
    out = dl.filter_pileup_adjacency_filter(d, methylation_threshold=0.7, adjacency_distance=1)
    assert out.position.tolist() == [1, 2, 3, 5, 7, 8, 9]
    d = dict(contig=np.array(["contig1"] * 5 + ["contig2"] * 5), position=list(range(5)) + list(range(5)),
             mod_type=np.array(["m6A", "5mC"] * 5), strand=np.array(["+"] * 5 + ["-"] * 5), fraction_mod=frac,
             Nvalid_cov=[10] * 10)
    out = dl.filter_pileup_adjacency_filter(d, methylation_threshold=0.7, adjacency_distance=1)
    assert out.position[out.contig == "contig1"].tolist() == [1, 2, 3]
    assert out.position[out.contig == "contig2"].tolist() == [0, 2, 3, 4]


def test_filters_random(dl):
    rng = np.random.default_rng(42)
    t = _table(rng)
    out = dl.filter_pileup(t)
    want = O.filter_pileup(t.Nvalid_cov)
    np.testing.assert_array_equal(out.position, t.position[want])
    assert 0 < want.sum() < len(want)

    out = dl.filter_pileup_minimummod_frequency(t)
    want = O.filter_pileup_minimummod_frequency(t.contig, t.mod_type, t.fraction_mod)
    np.testing.assert_array_equal(out.position, t.position[want])
    np.testing.assert_array_equal(out.contig, t.contig[want])
    assert 0 < want.sum() < len(want)

    for dist in (8, 1, 0):
        out = dl.filter_pileup_adjacency_filter(t, adjacency_distance=dist)
        want = O.filter_pileup_adjacency_filter(t.contig, t.strand, t.position, t.fraction_mod, 0.7, dist)
        np.testing.assert_array_equal(out.position, t.position[want])
        np.testing.assert_array_equal(out.fraction_mod, t.fraction_mod[want])
    # unsorted input gives the same set in input order
    perm = rng.permutation(len(t))
    shuffled = t.take(perm)
    out = dl.filter_pileup_adjacency_filter(shuffled)
    want = O.filter_pileup_adjacency_filter(shuffled.contig, shuffled.strand, shuffled.position, shuffled.fraction_mod)
    np.testing.assert_array_equal(out.position, shuffled.position[want])


def test_load_pileup_and_fasta(dl, tmp_path):
    rng = np.random.default_rng(3)
    t = _table(rng, n_contigs=2, length=3000)
    path = tmp_path / "pileup.bed"
    with open(path, "w") as f:
        for i in range(len(t)):
            pct = t.fraction_mod[i] * 100
            cov = int(t.Nvalid_cov[i])
            row = [t.contig[i], t.position[i], t.position[i] + 1, t.mod_type[i], cov, t.strand[i], t.position[i],
                   t.position[i] + 1, "255,0,0", cov, f"{pct:.2f}", int(round(pct * cov / 100)), 0, 0, 0, 0, 1, 0]
            f.write("\t".join(str(x) for x in row) + "\n")
    got = dl.load_pileup(str(path), with_counts=True)
    np.testing.assert_array_equal(got.position, t.position)
    np.testing.assert_array_equal(got.contig, t.contig)
    np.testing.assert_array_equal(got.mod_type.astype(str), t.mod_type.astype(str))  # "21839" stays a string
    np.testing.assert_array_equal(got.strand, t.strand)
    np.testing.assert_array_equal(got.Nvalid_cov, t.Nvalid_cov)
    want_frac = np.array([float(f"{v * 100:.2f}") for v in t.fraction_mod]) / 100  # dataload.py:85
    np.testing.assert_array_equal(got.fraction_mod, want_frac)
    assert got.extra["n_diff"].tolist() == [1] * len(t)
    fa = tmp_path / "a.fasta"
    fa.write_text(">c1 desc\nacgtNN\nACGT\n>c2\nTTTT\n")
    assert dl.load_fasta(str(fa)) == {"c1": "ACGTNNACGT", "c2": "TTTT"}
