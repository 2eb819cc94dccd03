#!/usr/bin/env python3
"""Host-side ceiling of the end-to-end path: pinned-host <-> device copy bandwidth of every GPU of the box, one at a
time and all at once (torchrun, one rank per GPU).  The e2e figure of bench.py moves 147 B device->host and 24 B
host->device per env-step; at N GPUs its aggregate rate cannot exceed what this prints.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_bw.py [--mbytes 512]
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def copy_gbs(dst, src, reps, dev):
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dst.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize(dev)
        best = max(best, src.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbytes", type=int, default=512)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n = a.mbytes << 20
    host, devb = torch.empty(n, dtype=torch.uint8, pin_memory=True), torch.empty(n, dtype=torch.uint8, device=dev)
    res = {}
    for name, (dst, src) in {"d2h": (host, devb), "h2d": (devb, host)}.items():
        alone = torch.zeros(world, dtype=torch.float64, device=dev)
        for r in range(world):                       # one GPU at a time
            if world > 1:
                dist.barrier()
            if r == rank:
                alone[r] = copy_gbs(dst, src, 5, dev)
        if world > 1:
            dist.all_reduce(alone)
            dist.barrier()
        # all GPUs at once: wall clock over 10 back-to-back copies per rank, started together
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        res[name] = {"alone_gbs_per_gpu": alone.tolist(), "together_aggregate_gbs": world * 10 * n / float(dt.item()) / 1e9,
                     "together_per_gpu_gbs": 10 * n / float(dt.item()) / 1e9}
    if rank == 0:
        out = {"gpus": world, "mbytes_per_copy": a.mbytes, "cpus_visible": len(os.sched_getaffinity(0)), **res}
        print(json.dumps(out))
        if a.out:
            json.dump(out, open(a.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
