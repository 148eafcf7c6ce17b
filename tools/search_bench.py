#!/usr/bin/env python
"""Lock-step motif search over many bins (scaled-down BASELINE cfg 3): every (bin, mod type) search advances together,
one K2 launch per score round and one K4 launch per expand / remove round (development tool).

    python tools/search_bench.py [--bins 16] [--bin-bp 1000000] [--cpu-bins 1]
"""
import argparse
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nanomotif_b200 as nmb  # noqa: E402
from nanomotif_b200 import growth, search, synth  # noqa: E402

PLANTED = [("GATC", 1), ("CTGCAG", 4), ("GA[AG]TC", 1), ("GCAC......GTT", 2), ("CAA..[AT]TG", 2), ("TTAA", 3), ("ACC.GT", 0),
           ("GAAG[CT]", 2)]


def build_bin(b, rng, bin_bp, depth):
    contigs, cols = {}, {k: [] for k in ("contig", "position", "strand", "fraction_mod")}
    n = int(rng.integers(2, 6))
    lens = np.maximum(20000, (rng.dirichlet(np.ones(n)) * bin_bp).astype(int))
    planted = [(PLANTED[i][0], PLANTED[i][1], "a") for i in rng.choice(len(PLANTED), size=int(rng.integers(1, 4)), replace=False)]
    for i, L in enumerate(lens):
        seq = synth.random_sequence(rng, int(L), float(rng.uniform(0.35, 0.65)), 2e-5)
        name = f"bin{b}_contig_{i}"
        contigs[name] = seq.tobytes().decode()
        p = synth.synth_pileup(seq, rng, depth=depth, mod_types=("a",), planted=planted)
        cols["contig"].append(np.full(len(p["position"]), name, dtype=object))
        cols["position"].append(p["position"])
        cols["strand"].append(np.where(p["strand"] == 0, "+", "-").astype(object))
        cols["fraction_mod"].append(p["fraction_mod"])
    return contigs, {k: np.concatenate(v) for k, v in cols.items()}, planted


def windows_for(asm, contigs, pile, pad, high):
    strand = (pile["strand"] == "-").astype(np.uint8)
    conf = pile["fraction_mod"] >= high
    ci, pos, st = [], [], []
    for name, seq in contigs.items():
        in_contig = conf & (pile["contig"] == name)
        for s in (0, 1):
            p = pile["position"][in_contig & (strand == s)]
            p = p[(p > pad) & (p < len(seq) - pad)]
            ci.append(np.full(len(p), asm.index[name]))
            pos.append(p)
            st.append(np.full(len(p), s, dtype=np.uint8))
    return growth.DeviceDNAarray.from_positions(asm, np.concatenate(ci), np.concatenate(pos), np.concatenate(st), pad)


def cfg3_main(args):
    """--cfg3: the lock-step search of EVERY (bin, mod type) of the cfg3 metagenome (bench.py's generator: 300 bins x
    5 Mbp, three mod types, planted motifs per bin), all on one GPU."""
    import bench
    from nanomotif_b200.dataload import DeviceRows
    from nanomotif_b200.device import DeviceAssembly, DevicePileup

    synth_m = bench.load_synth()
    plan = synth_m.cfg3_plan(bench.N_BINS + bench.SWEEP_EXTRA_BINS, bench.BIN_BP)
    dev = torch.device("cuda", 0)
    pad, low, high, min_kl, thr = 20, 0.3, 0.7, 0.05, 1.5
    bins = list(range(args.bins))
    t0 = time.perf_counter()
    names, lens, ranges, parts = [], [], {}, []
    for b in bins:
        d = synth_m.cfg3_bin_device(plan, b, dev, ascii_only=True)
        lo, hi = plan["ranges"][b]
        ranges[f"bin_{b}"] = (len(names), len(names) + hi - lo)
        names += [f"contig_{i}" for i in range(lo, hi)]
        lens.append(d["lengths"])
        parts.append(d["ascii"])
    lens = np.concatenate(lens)
    off = np.zeros(len(lens), dtype=np.int64)
    off[1:] = np.cumsum(lens)[:-1]
    asm = DeviceAssembly(names, lens, torch.cat(parts), off, dev)
    del parts
    pile = DevicePileup(asm, 3, low, high).clear()
    rows = []
    for b in bins:
        d = synth_m.cfg3_bin_device(plan, b, dev)
        cid = (d["contig"] + ranges[f"bin_{b}"][0]).to(torch.int32)
        pile.add_columns(cid, d["position"], d["strand"], d["fraction_mod"], d["mod_type"], sync=False)
        rows.append(DeviceRows(names, bench.MOD_TYPES, dev, contig_id=cid, position=d["position"], strand=d["strand"],
                               mod_type=d["mod_type"], fraction_mod=d["fraction_mod"]))
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    multi = nmb.MultiBinScorer.from_device(asm, pile, ranges, bench.MOD_TYPES, rows)
    n_rows = sum(len(r) for r in rows)
    print(f"cfg3: {len(bins)} bins, {asm.total_bp / 1e9:.2f} Gbp, {asm.n_contigs} contigs, {n_rows / 1e9:.2f} G pileup rows "
          f"(device generation + packing + class planes {t_gen:.1f} s)")
    searches, t_prep = [], 0.0
    pools = {}
    for mt in bench.MOD_TYPES:
        t0 = time.perf_counter()
        pool, pssms, totals = growth.prepare_searches(multi, mt, pad, high, seeds=[1 + b for b in bins])
        torch.cuda.synchronize()
        t_prep += time.perf_counter() - t0
        pools[mt] = pool
        for slot, b in enumerate(bins):
            co = search.find_candidates(mt, pad, pssms[slot], totals[slot], min_kl=min_kl, score_threshold=thr)
            searches.append((co, search.PoolBackend(multi.context(f"bin_{b}", mt), pool, slot), b, mt))
    print(f"windows + background PSSMs of all {len(searches)} searches (growth.prepare_searches, host random.sample stream kept) "
          f"{t_prep:.1f} s; windows: {sum(int(p.n_total) for p in pools.values())}")
    counts = {"score": 0, "expand": 0, "remove": 0, "motifs": 0}

    def counting(kind, fn):
        def hook(reqs):
            counts[kind] += 1
            if kind == "score":
                counts["motifs"] += sum(len(m) for _, m in reqs)
            return fn(reqs)
        return hook

    t0 = time.perf_counter()
    results = search.run_lockstep([(co, be) for co, be, _, _ in searches], counting("score", search.gpu_batch_score),
                                  counting("expand", search.gpu_batch_expand),
                                  counting("remove", search.gpu_batch_remove))
    torch.cuda.synchronize()
    t_search = time.perf_counter() - t0
    found = total = 0
    for (co, be, b, mt), res in zip(searches, results):
        best = set() if res is None else {m.string.strip(".") for m in res[1]}
        for motif, _, mtype in plan["planted"][b]:
            if mtype != mt:
                continue
            total += 1
            out = [""]
            for t in nmb.motif.tokenize(motif):
                out = [o + c for o in out for c in (t[1:-1] if t.startswith("[") else t)]
            found += motif in best or all(e in best for e in out)
    print(f"lock-step search of {len(searches)} (bin, mod type) pairs: {t_search:.1f} s; rounds: score {counts['score']} "
          f"({counts['motifs']} motifs), expand {counts['expand']}, remove {counts['remove']}")
    print(f"planted motifs recovered: {found} / {total}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg3", action="store_true", help="the whole cfg3 metagenome (device-generated), all mod types")
    ap.add_argument("--bins", type=int, default=16)
    ap.add_argument("--bin-bp", type=int, default=1_000_000)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--cpu-bins", type=int, default=1, help="bins also searched with the CPU oracle backend (slow)")
    ap.add_argument("--check-setup", action="store_true", help="compare the batched setup with the per-bin host-driven one")
    args = ap.parse_args()
    if args.cfg3:
        if args.bins == 16:
            args.bins = 300
        return cfg3_main(args)
    pad, low, high, min_kl, thr = 20, 0.3, 0.7, 0.05, 1.5
    rng = np.random.default_rng(5)
    t0 = time.perf_counter()
    bins, piles, truth = {}, [], {}
    for b in range(args.bins):
        contigs, pile, planted = build_bin(b, rng, args.bin_bp, args.depth)
        pile["mod_type"] = np.full(len(pile["position"]), "a", dtype=object)
        bins[f"bin{b}"], truth[f"bin{b}"] = contigs, planted
        piles.append(pile)
    pile = {k: np.concatenate([p[k] for p in piles]) for k in piles[0]}
    import pyarrow as pa

    # the input a polars frame would hold: Arrow columns (built here, outside the timed setup: it is the INPUT)
    table = pa.table({"contig": pa.array(pile["contig"], type=pa.large_string()), "position": pa.array(pile["position"]),
                      "strand": pa.array(pile["strand"], type=pa.large_string()),
                      "mod_type": pa.array(pile["mod_type"], type=pa.large_string()),
                      "fraction_mod": pa.array(pile["fraction_mod"])})
    t_gen = time.perf_counter() - t0
    nmb.MultiBinScorer(table.slice(0, 1000), {"w": {"w": "ACGT" * 100}}, ["a"], low, high)  # CUDA context, caches
    from nanomotif_b200 import device as _D

    _D._stager(torch.device("cuda", torch.cuda.current_device()))  # the pinned staging slots exist once per process
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    multi = nmb.MultiBinScorer(table, bins, ["a"], low, high)
    torch.cuda.synchronize()
    t_ingest = time.perf_counter() - t0
    # windows of every bin + every bin's background PSSM, batched on the device (growth.prepare_searches)
    pool, pssms, totals = growth.prepare_searches(multi, "a", pad, high, seeds=[1 + b for b in range(args.bins)])
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    if args.check_setup:  # the old per-bin host-driven setup must give the same windows and backgrounds
        asm = multi.assembly
        for b in range(args.bins):
            random.seed(1 + b)
            w = windows_for(asm, bins[f"bin{b}"], piles[b], pad, high)
            bg = growth.background_pssm(asm, bins[f"bin{b}"], "A", pad)
            assert torch.equal(pool.windows[pool.begin[b]:pool.end[b]], w.windows) and np.array_equal(bg, pssms[b])
    setups = list(zip(totals, pssms))
    counts = {"score": 0, "expand": 0, "remove": 0, "motifs": 0}

    def counting(kind, fn):
        def hook(reqs):
            counts[kind] += 1
            if kind == "score":
                counts["motifs"] += sum(len(m) for _, m in reqs)
            return fn(reqs)
        return hook

    searches = []
    for b, (total, bg) in enumerate(setups):
        co = search.find_candidates("a", pad, bg, total, min_kl=min_kl, score_threshold=thr)
        searches.append((co, search.PoolBackend(multi.context(f"bin{b}", "a"), pool, b)))
    t0 = time.perf_counter()
    results = search.run_lockstep(searches, counting("score", search.gpu_batch_score), counting("expand", search.gpu_batch_expand),
                                  counting("remove", search.gpu_batch_remove))
    torch.cuda.synchronize()
    t_search = time.perf_counter() - t0
    def expansions(m):  # concrete strings of a planted motif: the search reports bracket classes as separate motifs
        out = [""]
        for t in nmb.motif.tokenize(m):
            out = [o + c for o in out for c in (t[1:-1] if t.startswith("[") else t)]
        return out

    found = 0
    for b, res in enumerate(results):
        best = set() if res is None else {m.string.strip(".") for m in res[1]}
        found += all(p[0] in best or all(e in best for e in expansions(p[0])) for p in truth[f"bin{b}"])
    total_bp = sum(len(s) for cs in bins.values() for s in cs.values())
    print(f"{args.bins} bins, {total_bp / 1e6:.1f} Mbp, {len(pile['position']) / 1e6:.1f} M pileup rows (host generation {t_gen:.1f} s)")
    print(f"device setup from the Arrow table (ingest {t_ingest:.3f} s + windows / backgrounds of all bins) {t_setup:7.3f} s")
    print(f"lock-step search of {args.bins} (bin, mod type) pairs          {t_search:7.2f} s   rounds: score {counts['score']} "
          f"({counts['motifs']} motifs), expand {counts['expand']}, remove {counts['remove']}")
    print(f"bins whose planted motifs were all recovered: {found} / {args.bins}")
    if args.cpu_bins:
        from search_common import OracleBackend

        t0 = time.perf_counter()
        for b in range(args.cpu_bins):
            spec = dict(padding=pad, high=high, low=low, mod_type="a", random_seed=1 + b)
            be = OracleBackend(bins[f"bin{b}"], piles[b], spec)
            co = search.find_candidates("a", pad, be.bin_pssm, be.arr.shape[0], min_kl=min_kl, score_threshold=thr)
            search.run(co, be)
        t_cpu = (time.perf_counter() - t0) / args.cpu_bins
        print(f"same search with the CPU oracle backend (regex + np.isin, one core): {t_cpu:7.2f} s per bin "
              f"-> {t_cpu * args.bins:.0f} s for {args.bins} bins sequentially; lock-step GPU speed-up {t_cpu * args.bins / t_search:.0f}x")


if __name__ == "__main__":
    main()
