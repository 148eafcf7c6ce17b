#!/usr/bin/env python
"""Where the N-GPU step time goes (development tool): per-rank scan time, NCCL all-reduce alone, both."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
seq, pile, work = bench.build_cfg2(1 + (rank + int(os.environ.get("SEED_SHIFT", "0"))) % world)
state = bench.Cfg2Device(seq, pile, work, dev)


def timed(fn, k=30):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def scan_only():
    state.step()


def scan_allreduce():
    dist.all_reduce(state.step())


def allreduce_only():
    dist.all_reduce(state.out)


res = torch.tensor([timed(scan_only), timed(scan_allreduce), timed(allreduce_only, 100)], dtype=torch.float64, device=dev)
allr = [torch.zeros_like(res) for _ in range(world)]
dist.all_gather(allr, res)
if rank == 0:
    t = torch.stack(allr).cpu().numpy()
    print("rank  scan_ms  scan+allreduce_ms  allreduce_ms")
    for r in range(world):
        print(f"{r:4d}  {t[r,0]:7.3f}  {t[r,1]:7.3f}            {t[r,2]:7.3f}")
    print("max  ", t.max(axis=0), " mean ", t.mean(axis=0))
dist.destroy_process_group()
