#!/usr/bin/env python3
"""Host -> device path of the e2e leg with N ranks on one node: every rank uploads a clip-sized pinned buffer (1.25 GB), first
one rank at a time, then all ranks at once -- the e2e loss at N = 8 is the shared host side (memory controllers / PCIe root
ports of the sockets), not the GPUs.  Also prints where each rank's process and GPU sit (NUMA node, CPU affinity).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 profiles/exp_h2d_ranks.py"""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

NBYTES = 1253376000


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    host = torch.empty(NBYTES, dtype=torch.uint8, pin_memory=True)
    host.fill_(rank + 1)
    dev = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")
    back = torch.empty(76115285, dtype=torch.uint8, pin_memory=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def copy_ms(reps=3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
            back.copy_(dev[:back.numel()], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    solo = 0.0
    for r in range(world):       # one rank at a time
        sync_all()
        if r == rank:
            solo = copy_ms()
    sync_all()
    together = copy_ms()         # all ranks at once
    sync_all()
    try:
        aff = sorted(os.sched_getaffinity(0))
        aff_s = f"{aff[0]}-{aff[-1]} ({len(aff)} cpus)"
    except Exception:
        aff_s = "?"
    numa = "?"
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True)
        bid = q.stdout.strip().lower()
        if bid:
            path = f"/sys/bus/pci/devices/{bid[4:] if bid.startswith('0000') and len(bid) > 12 else bid}/numa_node"
            if os.path.exists(path):
                numa = open(path).read().strip()
    except Exception:
        pass
    rec = {"rank": rank, "solo_ms": round(solo, 2), "solo_gbs": round(NBYTES / solo / 1e6, 1), "together_ms": round(together, 2),
           "together_gbs": round(NBYTES / together / 1e6, 1), "cpu_affinity": aff_s, "gpu_numa_node": numa}
    allr = [None] * world
    if world > 1:
        dist.all_gather_object(allr, rec)
    else:
        allr = [rec]
    if rank == 0:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        print(json.dumps({"bytes_per_rank": NBYTES, "world": world, "ranks": allr,
                          "aggregate_together_gbs": round(sum(NBYTES / r["together_ms"] / 1e6 for r in allr), 1),
                          "aggregate_solo_gbs": round(sum(r["solo_gbs"] for r in allr), 1)}), flush=True)
        print(topo, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
