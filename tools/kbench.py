#!/usr/bin/env python
"""Kernel micro-benchmark for scan_count (development tool; not the judged bench).

    python tools/kbench.py [--bp 1500000000] [--contigs 17000] [--reps 20] [--grid 0] [--mpi 1]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nanomotif_b200 as nmb  # noqa: E402
from nanomotif_b200 import _lib  # noqa: E402
from nanomotif_b200.device import MotifPrograms, make_jobs, scan_count  # noqa: E402

MOTIFS = {"A": ("A", 0), "GATC": ("GATC", 1), "CCWGG": ("CC[AT]GG", 1), "GRNGAAGY": ("G[AG].GAAG[CT]", 5),
          "GCACN6GTT": ("GCAC......GTT", 2), "L13": ("ACGTTGCAAGCTA", 3)}


def build(device, total_bp, n_contigs, seed=3):
    from nanomotif_b200 import synth

    return synth.device_workload(device, total_bp, n_contigs, seed)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=1_500_000_000)
    ap.add_argument("--contigs", type=int, default=17000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--grid", type=int, nargs="*", default=[0])
    ap.add_argument("--mpi", type=int, nargs="*", default=[1])
    ap.add_argument("--motifs", nargs="*", default=["GATC", "GRNGAAGY", "A", "L13"])
    ap.add_argument("--batch", type=int, default=1, help="replicate each motif this many times in the launch")
    ap.add_argument("--random", type=int, default=0, help="also time a batch of N seeded random motifs (cfg2 work list)")
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    asm, pile = build(device, args.bp, args.contigs)
    rec_bytes = asm.n_tiles * (_lib.SEQ_REC_WORDS + _lib.CLS_REC_WORDS) * 4
    print(json.dumps({"bp": asm.total_bp, "tiles": asm.n_tiles, "record_bytes": rec_bytes}))
    sets = [(name, [nmb.Motif(*MOTIFS[name])] * args.batch) for name in args.motifs]
    if args.random:
        from nanomotif_b200 import synth
        sets.append((f"random{args.random}", [nmb.Motif(s, p) for s, p in synth.random_motifs(np.random.default_rng(1001), args.random, "A")]))
    for name, motifs in sets:
        args.batch = len(motifs)
        progs = MotifPrograms(motifs, device)
        jobs = make_jobs(1)
        jobs["motif_count"], jobs["tile_count"] = args.batch, asm.n_tiles
        jobs["contig_end"], jobs["n_groups"] = asm.n_contigs, 1
        res = torch.zeros((args.batch, 4), dtype=torch.int64, device=device)
        for grid in args.grid:
            for mpi in args.mpi:
                for _ in range(3):
                    scan_count(asm, pile, progs, jobs, args.batch, motifs_per_item=mpi, grid_ctas=grid, out=res)
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
                for a, b in evs:
                    a.record()
                    scan_count(asm, pile, progs, jobs, args.batch, motifs_per_item=mpi, grid_ctas=grid, out=res)
                    b.record()
                torch.cuda.synchronize()
                ts = np.array([a.elapsed_time(b) for a, b in evs])
                ms = float(np.median(ts))
                units = asm.total_bp * args.batch
                print(json.dumps({"motif": name, "grid": grid, "mpi": mpi, "batch": args.batch, "ms_med": round(ms, 4),
                                  "ms_min": round(float(ts.min()), 4), "Tunits_s": round(units / ms / 1e9, 3),
                                  "alg_GBs": round(0.75 * units / ms / 1e6, 1),
                                  "rec_GBs": round(rec_bytes / ms / 1e6, 1) if args.batch == 1 else None}))


if __name__ == "__main__":
    main()
