#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its time (development tool)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nanomotif_b200.device import DeviceAssembly, DevicePileup, MotifPrograms, scan_count  # noqa: E402

dev = torch.device("cuda", 0)
seq, pile, work = bench.build_cfg2(1)
state = bench.Cfg2Device(seq, pile, work, dev)
n = len(pile["position"])
host = {"ascii": torch.from_numpy(seq.copy()).pin_memory(), "contig_id": torch.zeros(n, dtype=torch.int32).pin_memory(),
        "position": torch.from_numpy(pile["position"]).pin_memory(), "strand": torch.from_numpy(pile["strand"]).pin_memory(),
        "mod_type": torch.from_numpy(pile["mod_type"]).pin_memory(), "fraction_mod": torch.from_numpy(pile["fraction_mod"]).pin_memory()}


def t(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r


ms, asm = t(lambda: DeviceAssembly(["c"], [len(seq)], host["ascii"], [0], dev))
print(f"DeviceAssembly (H2D 4.6 MB + pack + sync)      {ms:7.3f} ms")
ms, _ = t(lambda: [host[k].to(dev, non_blocking=True) for k in ("contig_id", "position", "strand", "mod_type", "fraction_mod")])
print(f"H2D of pileup columns (152 MB pinned)           {ms:7.3f} ms")
cols = [host[k].to(dev) for k in ("contig_id", "position", "strand", "fraction_mod", "mod_type")]
ms, dp = t(lambda: DevicePileup.from_columns(asm, cols[0], cols[1], cols[2], cols[3], 0.3, 0.7, cols[4], n_modtypes=3))
print(f"DevicePileup from device columns (memset+kernel) {ms:7.3f} ms")
ms, dp = t(lambda: DevicePileup.from_columns(asm, host["contig_id"], host["position"], host["strand"], host["fraction_mod"], 0.3, 0.7, host["mod_type"], n_modtypes=3))
print(f"DevicePileup from pinned host columns            {ms:7.3f} ms")
ms, progs = t(lambda: MotifPrograms(state.packed, dev))
print(f"MotifPrograms (H2D 192 KB + compile)             {ms:7.3f} ms")
ms, out = t(lambda: scan_count(asm, dp, progs, state.jobs, len(state.packed)))
print(f"scan_count (jobs upload + launch)                {ms:7.3f} ms")
ms, _ = t(lambda: out.cpu())
print(f"D2H counts                                       {ms:7.3f} ms")
hostd = dict(host, length=len(seq), packed=state.packed, jobs=state.jobs)
ms, _ = t(lambda: bench.e2e_step(hostd, dev))
print(f"e2e_step total                                   {ms:7.3f} ms")

# ---- streamed path (pipeline.score_host_blocks): host-side timeline of one step, ms since the call ----
from nanomotif_b200.device import compact_rows  # noqa: E402
from nanomotif_b200.pipeline import HostBlock, blocks_by_modtype, blocks_by_position, score_host_blocks  # noqa: E402

rows = compact_rows(np.zeros(n, np.int32), pile["position"], pile["strand"], pile["fraction_mod"], pile["mod_type"], 1)
pin = lambda bs: [HostBlock(*(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in b[:4]), b.modtypes, b.tiles) for b in bs]
args4 = (rows["position"], rows["flags"], rows["percent_x100"], rows["contig_row_off"])
styles = {"one block per mod type": pin(blocks_by_modtype(*args4, 3)),
          "tile ranges 1/8 + 3 x 7/24": pin(blocks_by_position(*args4, [len(seq)], 3)),
          "tile ranges 1/4 x 4": pin(blocks_by_position(*args4, [len(seq)], 3, (0.25, 0.25, 0.25, 0.25))),
          "tile ranges 1/6 + 5/12 x 2": pin(blocks_by_position(*args4, [len(seq)], 3, (1 / 6, 5 / 12, 5 / 12)))}
jobs0 = state.jobs.copy()
jobs0["tile_count"] = 0
out_host = torch.empty((len(state.packed), 4), dtype=torch.int64).pin_memory()
for name, blocks in styles.items():
    done = []
    for rep in range(8):
        tl = {}
        torch.cuda.synchronize()
        score_host_blocks(["c"], [len(seq)], host["ascii"], [0], blocks, state.packed, jobs0, len(state.packed), n_modtypes=3,
                          device=dev, out_host=out_host, timeline=tl)
        done.append(tl["done"])
    print(f"streamed step, {name}: median {np.median(done[2:]):.3f} ms; last timeline", {k: round(v, 3) for k, v in tl.items()})
