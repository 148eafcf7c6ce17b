#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.json by RUNNING the real reference.

Run in a container where /root/reference is mounted:

    python tests/golden/generate_golden.py

The reference (MicrobialDarkMatter/nanomotif 1.1.2) is imported unchanged through oracle/ref_shim.py
(stubs for the third-party packages missing from this image).  Every vector below is the output of a
reference function on a small seeded input; the inputs are stored with the outputs so that the tests
need neither the reference nor this script.  Functions that need polars (motif_model_contig/bin,
dataload filters) cannot run here and are not part of this file.
"""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402

nm = load_reference()
Motif = nm.motif.Motif
out = {"reference_version": getattr(nm, "__version__", "1.1.2"), "numpy": np.__version__}


def rand_seq(rng, n, n_rate=0.0):
    s = rng.choice(list("ACGT"), size=n)
    if n_rate:
        s[rng.random(n) < n_rate] = "N"
    return "".join(s)


rng = np.random.default_rng(20261017)

# ---- a1: subseq_indices ------------------------------------------------------------------------
cases = [("AATTAAATTAAGTAAAT", m) for m in ("AATT", "AA.T")]  # reference KAT inputs (tests/test_fasta.py:95-109)
cases += [("ACGTNACGTRACGTACNT", m) for m in ("ACGT", "AC.T", "A[CG]GT", "N", ".", "..", "A.")]
seq_a = rand_seq(rng, 3000, 0.01)
cases += [(seq_a, m) for m in ("GATC", "A", "CC[AT]GG", "G[AG].GAAG[CT]", "GCAC......GTT", "[ACG]A[CT]", "T....A",
                               ".GATC.", "A..............................T")]
out["subseq_indices"] = [dict(seq=s, motif=m, result=nm.utils.subseq_indices(m, s).tolist()) for s, m in cases]

# ---- a2: methylated_motif_occourances ------------------------------------------------------------
mm = []
for motif, pos in (("ACG", 0), ("GATC", 1), ("CC[AT]GG", 1), ("G[AG].GAAG[CT]", 5), ("A", 0)):
    for _ in range(2):
        sites = np.sort(rng.choice(len(seq_a), size=600, replace=False))
        meth, non = sites[:250], sites[250:]
        rng.shuffle(non)  # "pileup order" need not be sorted
        r = nm.find_motifs_bin.methylated_motif_occourances(Motif(motif, pos), seq_a, meth, non)
        mm.append(dict(seq="seq_a", motif=motif, mod_pos=pos, meth=meth.tolist(), nonmeth=non.tolist(),
                       result=[r[0].tolist(), r[1].tolist()]))
mm.append(dict(seq="TACGGACGCCACG", motif="ACG", mod_pos=0, meth=[1, 5], nonmeth=[10],
               result=[x.tolist() for x in nm.find_motifs_bin.methylated_motif_occourances(
                   Motif("ACG", 0), "TACGGACGCCACG", np.array([1, 5]), np.array([10]))]))
out["seq_a"] = seq_a
out["methylated_motif_occourances"] = mm

# ---- Motif helpers -----------------------------------------------------------------------------
mo = []
for s, p in (("....G[AG].GAAG[CT]....", 9), ("GATC", 1), ("..A..", 2), ("A[CT].[ACG]T", 0), ("CC[AT]GG", 1),
             (".....", 2), ("GCAC......GTT", 2), ("[AG]", 0)):
    m = Motif(s, p)
    st = m.new_stripped_motif()
    rc = st.reverse_compliment()
    mo.append(dict(motif=s, mod_pos=p, stripped=[st.string, st.mod_position], rc_of_stripped=[rc.string, rc.mod_position],
                   one_hot=m.one_hot().tolist(), split=m.split(), length=m.length(), iupac=m.iupac()))
out["motif"] = mo
out["from_iupac"] = [dict(iupac=s, regex=Motif(s, 0).from_iupac().string) for s in ("GATC", "CCWGG", "GRNGAAGY", "ACNNNNNVT", "BDHK")]

# ---- a5-a7: model and scores ---------------------------------------------------------------------
B = nm.model.BetaBernoulliModel


def model(n_mod, n_nomod):
    m = B()
    m.update(n_mod, n_nomod)
    return m


pairs = [((100, 3), (150, 500)), ((679, 74), (1000, 40000)), ((0, 0), (0, 0)), ((10, 0), (10, 0)), ((5, 5), (50, 50)),
         ((38207, 42), (38300, 1200000)), ((1, 0), (2, 3)), ((75, 17), (900, 20000))]
sc = []
searcher = nm.find_motifs_bin.MotifSearcher.__new__(nm.find_motifs_bin.MotifSearcher)
for nxt, cur in pairs:
    a, b = model(*nxt), model(*cur)
    sc.append(dict(next=list(nxt), cur=list(cur), score=float(nm.find_motifs_bin.predictive_evaluation_score(a, b)),
                   priority=float(searcher._priority_function(a, b)), mean_next=float(a.mean()),
                   variance_next=float(a.variance()), std_next=float(a.standard_deviation())))
out["scores"] = sc

# ---- a9-a12: motif-growth step -------------------------------------------------------------------
seq_g = rand_seq(np.random.default_rng(7), 20000, 0.002)
D = nm.seq.DNAsequence(seq_g)
pad = 20
gatc = nm.utils.subseq_indices("GATC", seq_g) + 1
ctag_rev = nm.utils.subseq_indices("GATC", seq_g) + 2  # the '-' strand A of GATC sits under the T... use as '-' sites
plus = [int(i) for i in gatc] + [5, 20, 21, len(seq_g) - 21, len(seq_g) - 20, 10000]  # boundary cases of seq.py:186
minus = [int(i) for i in ctag_rev] + [20, 21, 19979, 19980]
wp = D.sample_at_indices(plus, pad)
wm = D.sample_at_indices(minus, pad).reverse_compliment()
windows = [s.sequence for s in wp.sequences] + [s.sequence for s in wm.sequences]
ES = nm.seq.EqualLengthDNASet([nm.seq.DNAsequence(w) for w in windows])
arr = ES.convert_to_DNAarray()
random.seed(11)
bgset = D.sample_n_subsequences_unique(2 * pad + 1, 200, "A")
bg_windows = [s.sequence for s in bgset.sequences]
bin_pssm = bgset.pssm()
grow = dict(seq=seq_g, padding=pad, plus=plus, minus=minus, windows=windows, one_hot_sum=arr.sum(axis=0).tolist(),
            pssm_all=arr.pssm().tolist(), exact_pssm_all=ES.pssm().tolist(), bg_seed=11, bg_n=200, bg_base="A",
            bg_windows=bg_windows, bin_pssm=bin_pssm.tolist(), steps=[])
searcher.min_kl = 0.05
searcher.freq_threshold = 0.15
motif = Motif("." * pad + "A" + "." * pad, pad)
for _ in range(4):
    active = arr.copy().filter_sequence_matches(motif.one_hot())
    if active is None:
        break
    meth_pssm = active.pssm()
    from scipy.stats import entropy

    kl = entropy(meth_pssm, bin_pssm)
    children = list(searcher._motif_child_nodes_kl_dist_max(motif, meth_pssm, bin_pssm))
    removed = arr.filter_sequence_matches(motif.one_hot(), keep_matches=False)
    grow["steps"].append(dict(motif=motif.string, mod_pos=motif.mod_position, n_active=int(active.shape[0]),
                              column_counts=np.asarray(active.sum(axis=0)).tolist(), pssm=meth_pssm.tolist(),
                              kl=[float(v) for v in kl], children=[[c.string, c.mod_position] for c in children],
                              n_removed_rest=0 if removed is None else int(removed.shape[0])))
    if not children:
        break
    motif = children[0]
out["growth"] = grow

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
with open(path, "w") as f:
    json.dump(out, f)
print("wrote", path, os.path.getsize(path), "bytes")
