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
