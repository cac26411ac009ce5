#!/usr/bin/env python3
"""BASELINE.json configs[4] on the whole box: 3840x2160, 64 independent GOPs of 8 frames, i=16, r=64 full search, 4
references, GOP-sharded over the ranks of one node (GOP g -> rank g mod world, no collective on the data path; the per-GOP
container fragments are gathered on rank 0 and concatenated in GOP order).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 profiles/run_c4_8gpu.py

Prints one JSON line on rank 0: frames/s of the whole job (max over ranks of the encode time, host buffers in and out)."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import basic_video_codec_b200 as bvc  # noqa: E402
from basic_video_codec_b200.sharding import assign_gops, split_container_by_gop  # noqa: E402
from tests import synth  # noqa: E402

W, H, BS, R, QP, IP, NREF, NGOP = 3840, 2160, 16, 64, 4, 8, 4, 64


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mine = assign_gops(NGOP, world)[rank]
    # every GOP is its own seeded clip (independent content, like independent GOPs of a long sequence)
    buf = torch.empty((len(mine) * IP, H, W), dtype=torch.uint8, pin_memory=True)
    frames = buf.numpy()
    for i, g in enumerate(mine):
        frames[i * IP:(i + 1) * IP] = synth.moving_clip(2160 + g, H, W, IP, step=3, clamp=48, noise=2)
    out_t = torch.empty(frames.nbytes // 2, dtype=torch.uint8, pin_memory=True)
    out = out_t.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with bvc.Context(W, H, BS, R, QP, NREF, False, False, IP, device=local, max_lanes=len(mine)) as ctx:
        ctx.encode_clip_into(frames, out)     # warm-up
        barrier()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            ln = ctx.encode_clip_into(frames, out)
        barrier()
        dt = (time.perf_counter() - t0) / reps
        _, clip_ms = ctx.last_kernel_times()
        ctx.set_lane_groups(1)
        ctx.encode_clip_into(frames, out)
        kt, _ = ctx.last_kernel_times()
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt_max = float(t.item())
    parts = list(zip(mine, split_container_by_gop(out[:ln].tobytes(), [IP] * len(mine))))
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, [(g, hashlib.sha256(p).hexdigest(), len(p)) for g, p in parts])
    else:
        gathered = [[(g, hashlib.sha256(p).hexdigest(), len(p)) for g, p in parts]]
    if rank == 0:
        allp = sorted(x for lst in gathered for x in lst)
        assert [g for g, _, _ in allp] == list(range(NGOP))
        print(json.dumps({
            "workload": "BASELINE configs[4]: 3840x2160, 64 GOPs x 8 frames, i=16, r=64 full search, 4 refs, QP 4",
            "n_gpus": world, "gops_per_rank": len(mine), "frames": NGOP * IP, "seconds": dt_max, "frames_per_s": NGOP * IP / dt_max,
            "rank0_device_ms": clip_ms, "rank0_kernel_ms": {k: v[0] for k, v in kt.items()},
            "note": "rank0_kernel_ms from an extra pass with one lane group (kernels serialised on one stream)",
            "container_bytes": sum(n for _, _, n in allp),
            "container_sha256_of_gop_hashes": hashlib.sha256("".join(h for _, h, _ in allp).encode()).hexdigest(),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
