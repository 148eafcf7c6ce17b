#!/usr/bin/env python
"""Generate tests/golden/motif_algebra.json by RUNNING the real reference's Motif and BetaBernoulliModel.

    python tests/golden/generate_motif_golden.py        (needs /root/reference)

The host side of the path mirrors two value types of the reference -- nanomotif.motif.Motif (nanomotif/motif.py:18-359)
and nanomotif.model.BetaBernoulliModel (nanomotif/model.py:11-126) -- because the search driver, the priority function
and the merge step consume them.  This file records what the REFERENCE types return on seeded random motifs / counts,
method by method, so that nanomotif_b200.motif.Motif and nanomotif_b200.model can be checked on a box without the
reference tree (tests/test_motif_algebra.py).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402

nm = load_reference()
Motif = nm.motif.Motif
rng = np.random.default_rng(20261018)
CLASSES = ["[AC]", "[AG]", "[AT]", "[CG]", "[CT]", "[GT]", "[ACG]", "[ACT]", "[AGT]", "[CGT]"]


def random_motif():
    """What the search and the merge step produce: bases, wildcards, sorted bracket classes; optional '.' flanks."""
    n = int(rng.integers(1, 13))
    toks = []
    for _ in range(n):
        r = rng.random()
        toks.append("." if r < 0.35 else (str(rng.choice(CLASSES)) if r < 0.5 else str(rng.choice(list("ACGT")))))
    lead, trail = (int(rng.integers(0, 4)) if rng.random() < 0.4 else 0 for _ in range(2))
    toks = ["."] * lead + toks + ["."] * trail
    return "".join(toks), int(rng.integers(0, len(toks)))


motifs = [random_motif() for _ in range(260)] + [("GATC", 1), ("....G[AG].GAAG[CT]....", 9), (".....", 2), ("A", 0),
                                                 ("..A..", 2), ("A.C", 0), ("A..C", 0), ("G.A.C", 2), ("[AG]", 0)]
single = []
for s, p in motifs:
    m = Motif(s, p)
    st = m.new_stripped_motif()
    rc = m.reverse_compliment()
    single.append(dict(
        motif=s, mod_pos=p, split=m.split(), length=m.length(), stripped=[st.string, st.mod_position],
        reverse_compliment=[rc.string, rc.mod_position], one_hot=m.one_hot().tolist(), iupac=m.iupac(),
        isolated=[int(m.count_isolated_bases(isolation_size=k)) for k in (1, 2, 3)], repr=repr(m), hash_equal=hash(m) == hash(Motif(s, p))))
pairs = []
for _ in range(1500):
    (s1, p1), (s2, p2) = motifs[int(rng.integers(len(motifs)))], motifs[int(rng.integers(len(motifs)))]
    if rng.random() < 0.5:  # related pairs: a child of the first motif (one wildcard filled in), as the search makes them
        toks = Motif(s1, p1).split()
        dots = [i for i, t in enumerate(toks) if t == "." and i != p1]
        if dots:
            toks[int(rng.choice(dots))] = str(rng.choice(list("ACGT")))
            s2, p2 = "".join(toks), p1
    a, b = Motif(s1, p1), Motif(s2, p2)
    pairs.append(dict(a=[s1, p1], b=[s2, p2], a_sub_b=bool(a.sub_motif_of(b)), b_sub_a=bool(b.sub_motif_of(a)),
                      eq=bool(a == b), ne=bool(a != b)))
groups = []
for _ in range(60):
    idx = rng.choice(len(motifs), size=int(rng.integers(1, 6)), replace=False)
    s, p = motifs[int(rng.integers(len(motifs)))]
    groups.append(dict(motif=[s, p], others=[list(motifs[int(i)]) for i in idx],
                       sub_motif_of_any=bool(Motif(s, p).sub_motif_of_any([Motif(*motifs[int(i)]) for i in idx]))))
iupac = []
for _ in range(80):
    s = "".join(rng.choice(list("ACGTRYSWKMBDHVN"), size=int(rng.integers(1, 12))))
    iupac.append(dict(iupac=s, regex=Motif(s, 0).from_iupac().string))

B = nm.model.BetaBernoulliModel
models = []
for _ in range(120):
    n_mod, n_nomod = (int(rng.integers(0, 10 ** int(rng.integers(1, 7)))) for _ in range(2))
    x, y = (int(rng.integers(0, 2000)) for _ in range(2))
    if rng.random() < 0.15:
        x = y = 0
    m = B()
    m.update(n_mod, n_nomod)
    rec = dict(n_mod=n_mod, n_nomod=n_nomod, alpha=m._alpha, beta=m._beta, raw=list(m.get_raw_counts()), mean=float(m.mean()),
               variance=float(m.variance()), std=float(m.standard_deviation()), x=x, y=y,
               posterior_predictive=float(m.posterior_predictive(x, y)),
               posterior_predictive_per_obs=float(m.posterior_predictive_per_obs(x, y)), state=m.__getstate__())
    m.reset()
    rec["after_reset"] = [m._alpha, m._beta]
    models.append(rec)
custom = B(2, 7)
custom.update(3, 4)
out = dict(reference_version=getattr(nm, "__version__", "1.1.2"), single=single, pairs=pairs, groups=groups, from_iupac=iupac,
           models=models, custom_prior=dict(alpha=custom._alpha, beta=custom._beta, raw=list(custom.get_raw_counts()),
                                            mean=float(custom.mean())))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "motif_algebra.json")
with open(path, "w") as f:
    json.dump(out, f)
print(path, os.path.getsize(path), "bytes;", len(single), "motifs,", len(pairs), "pairs")
