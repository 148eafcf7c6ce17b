#!/usr/bin/env python
"""cfg4 timing of the contig x motif methylation-pattern table (K5): 50 000 contigs x 500 motifs (development tool).

    python tools/pattern_bench.py [--contigs 50000] [--bp 1500000000] [--motifs 500]
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
from nanomotif_b200.pattern import PatternIndex, pattern_table  # noqa: E402

IUPAC = {"A": "A", "C": "C", "G": "G", "T": "T", "R": "[AG]", "Y": "[CT]", "S": "[CG]", "W": "[AT]", "K": "[GT]", "M": "[AC]",
         "B": "[CGT]", "D": "[AGT]", "H": "[ACT]", "V": "[ACG]", "N": "."}


def bin_motifs(rng, n):
    """IUPAC motifs in the style of bin-motifs.tsv: length 4-14, mostly ACGT, some degenerate letters and N gaps,
    modified base = an A of the motif."""
    out = []
    while len(out) < n:
        L = int(rng.integers(4, 15))
        letters = list(rng.choice(list("ACGT"), size=L))
        for _ in range(int(rng.integers(0, 3))):
            letters[int(rng.integers(1, L - 1))] = str(rng.choice(list("RYSWKMN")))
        if L >= 8 and rng.random() < 0.4:
            a = int(rng.integers(2, L - 5))
            for j in range(a, a + int(rng.integers(3, 6))):
                if j < L - 2:
                    letters[j] = "N"
        if "A" not in letters:
            letters[int(rng.integers(0, L))] = "A"
        if letters[0] == "N" or letters[-1] == "N":
            continue
        mp = int(rng.choice([i for i, c in enumerate(letters) if c == "A"]))
        out.append(nmb.Motif("".join(IUPAC[c] for c in letters), mp))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contigs", type=int, default=50_000)
    ap.add_argument("--bp", type=int, default=1_500_000_000)
    ap.add_argument("--motifs", type=int, default=500)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    asm, rows = synth.device_pattern_workload(dev, args.bp, args.contigs)
    n_rows = int(rows["position"].numel())
    motifs = bin_motifs(np.random.default_rng(2), args.motifs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index = PatternIndex(asm, rows)
    torch.cuda.synchronize()
    t_index = time.perf_counter() - t0
    print(f"assembly {asm.total_bp / 1e9:.2f} Gbp, {asm.n_contigs} contigs, {n_rows / 1e6:.0f} M rows ({index.n_valid_rows / 1e6:.0f} M valid)")
    print(f"index build                         {t_index * 1e3:9.1f} ms")
    # device time of the scan kernel alone (one batch, phase 0), the rest of pattern_table is copies and host work
    import ctypes as C

    from nanomotif_b200._lib import check, lib, ptr
    from nanomotif_b200.device import MotifPrograms, _stream

    chunk = motifs[:args.batch]
    progs = MotifPrograms(chunk, dev, strip=False)
    stats = torch.zeros((len(chunk) * asm.n_contigs, 3), dtype=torch.int64, device=dev)
    view = asm.view()
    from nanomotif_b200.device import _work_counter

    for mpi, balanced in ((32, True), (32, False), (16, True), (8, True)):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for rep in range(3):
            if rep == 1:
                ev[0].record()
            if balanced:
                check(lib.nmb_pattern_scan_balanced(C.byref(view), ptr(index.valid), ptr(index.rank_dir), ptr(index.payload),
                                                    ptr(progs.programs), len(chunk), mpi, progs.max_len, 0, ptr(stats), None, None,
                                                    None, 0, ptr(_work_counter(dev)), _stream()), "scan")
            else:
                check(lib.nmb_pattern_scan(C.byref(view), ptr(index.valid), ptr(index.rank_dir), ptr(index.payload),
                                           ptr(progs.programs), len(chunk), mpi, progs.max_len, 0, ptr(stats), None, None, None,
                                           0, _stream()), "scan")
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 2
        print(f"scan kernel, {len(chunk)} motifs, mpi {mpi}, {'dynamic' if balanced else 'static '} items: {ms:8.2f} ms  "
              f"{len(chunk) * asm.total_bp / ms / 1e9:.2f}e12 motif*bp/s")
    for median in (False, True):
        pattern_table(index, motifs[:args.batch], median, args.batch)  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        stats, _ = pattern_table(index, motifs, median, args.batch)
        dt = time.perf_counter() - t0
        units = len(motifs) * asm.total_bp
        print(f"{len(motifs)} motifs x {asm.n_contigs} contigs, {'median' if median else 'weighted mean'}: {dt * 1e3:9.1f} ms  "
              f"{units / dt / 1e12:.2f}e12 motif*bp/s  cells with observations {int((stats[:, :, 0] > 0).sum()) / 1e6:.1f} M  "
              f"observations {stats[:, :, 0].sum() / 1e6:.0f} M")


if __name__ == "__main__":
    main()
