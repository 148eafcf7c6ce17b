"""Parity of the motif-growth step (K4: windows, filter, PSSM, KL) against the CPU oracle."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import restate as O


@pytest.fixture(scope="module")
def env():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nanomotif_b200 as nmb
    from nanomotif_b200 import growth, synth
    from nanomotif_b200.device import DeviceAssembly

    rng = np.random.default_rng(7)
    contigs, piles = {}, []
    for i, L in enumerate((200000, 30000, 70001)):
        seq = synth.random_sequence(rng, L, 0.5, 5e-5 if i else 0.0)
        contigs[f"c{i}"] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=25, mod_types=("a",), planted=(("GATC", 1, "a"), ("CA....TG", 1, "a")))
        p["contig_id"] = np.full(len(p["position"]), i, dtype=np.int32)
        piles.append(p)
    pile = {k: np.concatenate([p[k] for p in piles]) for k in piles[0]}
    asm = DeviceAssembly.from_sequences(contigs)
    return dict(nmb=nmb, growth=growth, contigs=contigs, pile=pile, asm=asm)


def _oracle_windows(contigs, pile, high, padding):
    out = []
    for i, (name, seq) in enumerate(contigs.items()):
        sel = (pile["contig_id"] == i) & (pile["fraction_mod"] >= high)
        plus = pile["position"][sel & (pile["strand"] == 0)].tolist()
        minus = pile["position"][sel & (pile["strand"] == 1)].tolist()
        out += O.methylation_windows(seq, plus, minus, padding)
    return out


@pytest.mark.parametrize("padding", [20, 5, 30])
def test_windows_filter_pssm_kl(env, padding):
    g, pile, contigs = env["growth"], env["pile"], env["contigs"]
    width = 2 * padding + 1
    arr = g.methylation_windows(env["asm"], pile["contig_id"], pile["position"], pile["strand"], pile["fraction_mod"],
                                0.7, padding)
    want_windows = _oracle_windows(contigs, pile, 0.7, padding)
    assert arr.shape == (len(want_windows), width, 4)
    ref = O.one_hot_windows(want_windows)
    np.testing.assert_array_equal(arr.column_counts(), ref.sum(axis=0))  # exact integer histogram
    np.testing.assert_allclose(arr.pssm(), O.pssm(ref), rtol=0, atol=0)

    root = "." * padding + "A" + "." * padding
    bg = np.full((4, width), 0.25)
    bg[:, 3] = [0.4, 0.1, 0.2, 0.3]
    # grow along the planted GATC motif: root -> children, checking each expansion against the oracle
    motif = env["nmb"].Motif(root, padding)
    active_ref = ref
    for _ in range(3):
        mask = O.motif_one_hot(motif.string)
        sel, active_ref = O.filter_sequence_matches(ref, mask, True)
        active = arr.copy().filter_sequence_matches(motif.one_hot())
        assert active.shape[0] == active_ref.shape[0]
        np.testing.assert_array_equal(active.column_counts(), active_ref.sum(axis=0))
        meth = active.pssm()
        np.testing.assert_array_equal(meth, O.pssm(active_ref))
        kl_ref, children_ref = O.kl_children(motif.string, motif.mod_position, O.pssm(active_ref), bg)
        n_act, pssm_dev, kl_dev = arr.expand([motif], bg)
        assert n_act[0] == active_ref.shape[0]
        np.testing.assert_allclose(pssm_dev[0], O.pssm(active_ref), rtol=1e-12)
        np.testing.assert_allclose(kl_dev[0], kl_ref, rtol=1e-6, atol=1e-12)  # north-star tolerance
        children = g.kl_children(motif, meth, bg, kl=kl_dev[0])
        assert [(c.string, c.mod_position) for c in children] == children_ref
        if not children:
            break
        motif = children[0]
    # removal of matching rows (find_motifs_bin.py:803)
    mask = O.motif_one_hot(motif.string)
    _, rest_ref = O.filter_sequence_matches(ref, mask, False)
    rest = arr.filter_sequence_matches(motif.one_hot(), keep_matches=False)
    assert rest.shape[0] == rest_ref.shape[0]
    np.testing.assert_array_equal(rest.column_counts(), rest_ref.sum(axis=0))
    # chained: filter the remainder again with another motif
    m2 = env["nmb"].Motif("." * (padding - 1) + "CA" + "." * padding, padding)
    _, r2_ref = O.filter_sequence_matches(rest_ref, O.motif_one_hot(m2.string), True)
    r2 = rest.filter_sequence_matches(m2.one_hot())
    if r2_ref is None:
        assert r2 is None
    else:
        np.testing.assert_array_equal(r2.column_counts(), r2_ref.sum(axis=0))


def test_background_pssm(env):
    g, contigs = env["growth"], env["contigs"]
    padding = 20
    random.seed(1)
    got = g.background_pssm(env["asm"], contigs, "A", padding)
    rng = random.Random(1)
    windows = []
    for name, seq in contigs.items():
        windows += O.sample_background(seq, 2 * padding + 1, O.n_background_samples(len(seq)), "A", rng)
    want = O.background_pssm(windows)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-15)


def test_empty_filter_returns_none(env):
    g, pile = env["growth"], env["pile"]
    arr = g.methylation_windows(env["asm"], pile["contig_id"], pile["position"], pile["strand"], pile["fraction_mod"],
                                0.7, 20)
    impossible = np.zeros((41, 4), dtype=int)
    with pytest.warns(UserWarning):
        assert arr.filter_sequence_matches(impossible) is None


def test_prepare_searches_equals_the_per_bin_setup(env):
    """growth.prepare_searches (all bins at once, rows selected and ordered on the device, valid background starts
    enumerated on the device, the reference's random.sample stream kept on the host) gives the same windows and the
    same background PSSM as the per-bin host-driven setup -- and as the oracle's restatement of find_motifs_bin.py:625-686."""
    import torch

    nmb, g = env["nmb"], env["growth"]
    from nanomotif_b200 import synth

    rng = np.random.default_rng(17)
    bins, cols = {}, {k: [] for k in ("contig", "position", "strand", "mod_type", "fraction_mod")}
    for b, lens in (("b0", (60000, 9000)), ("b1", (30000,)), ("b2", (12000, 45000, 8000))):
        bins[b] = {}
        for i, L in enumerate(lens):
            name = f"{b}_{i}"
            seq = synth.random_sequence(rng, L, 0.5, 3e-5)
            bins[b][name] = seq.tobytes().decode()
            p = synth.synth_pileup(seq, rng, depth=20, mod_types=("a", "m"), planted=(("GATC", 1, "a"), ("CC[AT]GG", 1, "m")))
            n = len(p["position"])
            cols["contig"].append(np.full(n, name, dtype=object))
            cols["position"].append(p["position"])
            cols["strand"].append(np.where(p["strand"] == 0, "+", "-").astype(object))
            cols["mod_type"].append(np.array(["a", "m"], dtype=object)[p["mod_type"]])
            cols["fraction_mod"].append(p["fraction_mod"])
    pile = {k: np.concatenate(v) for k, v in cols.items()}
    multi = nmb.MultiBinScorer(pile, bins, ["a", "m"], 0.3, 0.7)
    asm = multi.assembly
    for mod_type, base in (("a", "A"), ("m", "C")):
        for pad in (20, 7):
            pool, pssms, totals = g.prepare_searches(multi, mod_type, pad, 0.7, seeds=[11, 12, 13])
            sel_mt = pile["mod_type"] == mod_type
            for slot, (b, contigs) in enumerate(bins.items()):
                sel = sel_mt & np.isin(pile["contig"], list(contigs))
                cid = np.array([asm.index[c] for c in pile["contig"][sel]], dtype=np.int32)
                want = g.methylation_windows(asm, cid, pile["position"][sel], (pile["strand"][sel] == "-").astype(np.uint8),
                                             pile["fraction_mod"][sel], 0.7, pad)
                got = pool.windows[pool.begin[slot]:pool.end[slot]]
                assert totals[slot] == want.n_total == int(got.shape[0])
                assert torch.equal(got, want.windows)
                random.seed(11 + slot)
                np.testing.assert_array_equal(pssms[slot], g.background_pssm(asm, contigs, base, pad))
                # and the oracle: windows around the same sites, background from the same random stream
                rng_o = random.Random(11 + slot)
                bg = []
                for seq in contigs.values():
                    bg += O.sample_background(seq, 2 * pad + 1, O.n_background_samples(len(seq)), base, rng_o)
                np.testing.assert_array_equal(pssms[slot], O.background_pssm(bg))
