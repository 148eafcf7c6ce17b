#!/usr/bin/env python
"""BASELINE cfg 5 timing: every IUPAC motif of length 4..8 at every modified position over a cfg3/cfg5-shaped assembly
(development tool).  Also cross-checks a sample of table entries against K2 scans of the same motifs.

    python tools/sweep_bench.py [--bp 2000000000] [--contigs 20000]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nanomotif_b200 as nmb  # noqa: E402
from nanomotif_b200 import synth  # noqa: E402
from nanomotif_b200.device import MotifPrograms, make_jobs, scan_count  # noqa: E402
from nanomotif_b200.sweep import IUPAC_ORDER, SweepIndex  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=2_000_000_000)
    ap.add_argument("--contigs", type=int, default=20_000)
    ap.add_argument("--check", type=int, default=64, help="table entries cross-checked against K2")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    asm, pile = synth.device_workload(dev, args.bp, args.contigs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index = SweepIndex(asm, pile, 0).add()
    torch.cuda.synchronize()
    t_hist = time.perf_counter() - t0
    print(f"assembly {asm.total_bp / 1e9:.2f} Gbp, {asm.n_contigs} contigs; histogram pass {t_hist * 1e3:.1f} ms "
          f"({asm.total_bp / t_hist / 1e9:.1f} Gbp/s), {int(index.raw.long().sum().item()) / 1e9:.2f} G increments")
    n_motifs, t_tab, t_filt, n_cand = 0, 0.0, 0.0, 0
    for k in range(4, 9):
        for o in range(k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n_mod, n_nomod = index.table(k, o, "A")
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            cand = index.candidates(k, o, "A", min_mean=0.6, min_mod=1000) if k <= 6 else []
            torch.cuda.synchronize()
            t_tab += t1 - t0
            t_filt += time.perf_counter() - t1
            n_motifs += int(n_mod.numel())
            n_cand += len(cand)
            del n_mod, n_nomod
    units = n_motifs * asm.total_bp
    print(f"all tables k = 4..8, every modified position: {n_motifs / 1e9:.3f} G motifs in {t_tab * 1e3:.0f} ms "
          f"(+ filters for k <= 6: {t_filt * 1e3:.0f} ms, {n_cand} candidates)")
    print(f"equivalent brute-force work {units:.2e} motif*bp -> {units / (t_hist + t_tab):.1e} motif*bp/s equivalent; "
          f"at K2's 1e13 motif*bp/s the same table would take {units / 1e13 / 3600:.1f} GPU-hours")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index.add_bipartite()
    torch.cuda.synchronize()
    t_bip = time.perf_counter() - t0
    n_bip = sum(4 ** (a + b) * (a + b) for a in (3, 4) for b in (3, 4)) * 5
    print(f"bipartite X{{3,4}} N{{4..8}} Y{{3,4}}: {n_bip / 1e6:.2f} M (motif, position) pairs in {t_bip * 1e3:.0f} ms, "
          f"{int(index.bip_raw.long().sum().item()) / 1e9:.1f} G increments (K2 brute force: {n_bip * asm.total_bp / 1e13:.0f} s)")
    # cross-check against K2
    rng = np.random.default_rng(3)
    motifs, want = [], []
    while len(motifs) < args.check:
        k = int(rng.integers(4, 9))
        o = int(rng.integers(0, k))
        s = list(rng.choice(list(IUPAC_ORDER), size=k, p=[0.12] * 4 + [0.03] * 10 + [0.22]))
        s[o] = "A"
        s = "".join(s)
        if s[0] == "N" or s[-1] == "N":
            continue
        motifs.append(nmb.Motif(s, o).from_iupac())
        want.append(index.counts(s, o, keep=False))
    jobs = make_jobs(1)
    jobs["motif_count"], jobs["tile_count"], jobs["contig_end"], jobs["n_groups"] = len(motifs), asm.n_tiles, asm.n_contigs, 1
    c = scan_count(asm, pile, MotifPrograms(motifs, dev), jobs, len(motifs)).cpu().numpy()
    got = [(int(r[0] + r[2]), int(r[1] + r[3])) for r in c]
    assert got == want, [(m, g, w) for m, g, w in zip(motifs, got, want) if g != w][:5]
    print(f"{len(motifs)} sampled table entries equal K2 scans of the same motifs (max n_mod {max(w[0] for w in want)})")


if __name__ == "__main__":
    main()
