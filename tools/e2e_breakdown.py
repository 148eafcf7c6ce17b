#!/usr/bin/env python
"""Where the end-to-end step of bench.py goes (development tool): host table -> device, name resolution, class planes,
packing, 64 frontier rounds.  python tools/e2e_breakdown.py [--bins 8]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bins", type=int, default=8)
args = ap.parse_args()
synth = bench.load_synth()
plan = synth.cfg3_plan(400, 5_000_000)
bins = list(range(args.bins))
hb = bench.host_bins(synth, plan, bins)

import torch  # noqa: E402

import nanomotif_b200 as nmb  # noqa: E402
from nanomotif_b200 import dataload, device as D  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
table = bench.arrow_table(plan, hb)
bins_arg = {f"bin_{b}": bench.contig_strings(plan, b, hb[b]) for b in bins}
names = [n for cs in bins_arg.values() for n in cs]
print(f"{table.num_rows} rows, {bench.table_bytes(table) / 1e9:.2f} GB of table, {sum(map(len, bins_arg.values()))} contigs")


def t(label, fn, reps=3, nbytes=None):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    extra = f"  {nbytes / ms / 1e6:.1f} GB/s" if nbytes else ""
    print(f"{label:48s} {ms:9.2f} ms{extra}")
    return out


pos = table.column("position").chunk(0).to_numpy()
print("--- raw copies of the position column (%.0f MB) ---" % (pos.nbytes / 1e6))
t("torch pageable .to()", lambda: torch.from_numpy(pos).to(dev), nbytes=pos.nbytes)
pinned = torch.from_numpy(pos).pin_memory()
t("torch pinned .to(non_blocking)", lambda: pinned.to(dev, non_blocking=True), nbytes=pos.nbytes)
for n in (1, 2, 4, 8, 16):
    os.environ["NMB_STAGE_THREADS"] = str(n)
    D._STAGERS.clear()
    t(f"stager, {n} threads", lambda: D._to_device(pos, dev), nbytes=pos.nbytes)
del os.environ["NMB_STAGE_THREADS"]
D._STAGERS.clear()
print("--- stages ---")
rows = t("rows_from_table (H2D + 3 name lookups)", lambda: dataload.rows_from_table(table, names, bench.MOD_TYPES, dev, {}, with_coverage=False),
         nbytes=bench.table_bytes(table))
asm = t("DeviceAssembly.from_sequences (join + H2D + pack)", lambda: D.DeviceAssembly.from_sequences({n: s for cs in bins_arg.values() for n, s in cs.items()}, dev))
pile = D.DevicePileup(asm, 3, 0.3, 0.7)
t("class planes (clear + add)", lambda: pile.clear().add_columns(rows.contig_id, rows.position, rows.strand, rows.fraction_mod, rows.mod_type))
scorer = t("MultiBinScorer(table, bins) whole constructor", lambda: nmb.MultiBinScorer(table, bins_arg, bench.MOD_TYPES, 0.3, 0.7, dev), reps=2)
wl = bench.job_worklists(synth, bins)
keys = sorted(wl)
motifs = {k: [[nmb.Motif(m, p) for m, p in kids] for kids in wl[k]] for k in keys}
ctx = {k: scorer.context(f"bin_{k[0]}", bench.MOD_TYPES[k[1]]) for k in keys}
t("64 score_batch rounds (counts to the host each)", lambda: [scorer.score_batch([(ctx[k], motifs[k][r]) for k in keys]) for r in range(64)], reps=2)
