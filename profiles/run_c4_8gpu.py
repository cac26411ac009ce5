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


def measure(rank, world, local, reps=2):
    """The measurement itself; torch.distributed already initialised by the caller when world > 1 (bench.py calls this at
    N = 8 so that the configuration appears in the driver's scaling record).  Returns the JSON object on rank 0, else None."""
    mine = assign_gops(NGOP, world)[rank]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # set-up and warm-up without collectives; the ranks then agree that all of them got this far, so that a failure on one
    # rank (memory, a missing library) cannot leave the others waiting in a barrier
    ctx, err = None, None
    try:
        # every GOP is its own seeded clip (independent content, like independent GOPs of a long sequence)
        buf = torch.empty((len(mine) * IP, H, W), dtype=torch.uint8, pin_memory=True)
        frames = buf.numpy()
        for i, g in enumerate(mine):
            frames[i * IP:(i + 1) * IP] = synth.moving_clip(2160 + g, H, W, IP, step=3, clamp=48, noise=2)
        out_t = torch.empty(frames.nbytes // 2, dtype=torch.uint8, pin_memory=True)
        out = out_t.numpy()
        ctx = bvc.Context(W, H, BS, R, QP, NREF, False, False, IP, device=local, max_lanes=len(mine))
        ctx.encode_clip_into(frames, out)     # warm-up
    except Exception as e:
        err = repr(e)
    okt = torch.tensor([0 if err else 1], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if int(okt.item()) == 0:
        if ctx is not None:
            ctx.close()
        return {"error": err or "another rank failed during set-up"} if rank == 0 else None
    with ctx:
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            ln = ctx.encode_clip_into(frames, out)
        barrier()
        dt = (time.perf_counter() - t0) / reps
        _, clip_ms = ctx.last_kernel_times()
        ctx.set_lane_groups(1)
        ctx.encode_clip_into(frames, out)
        kt, _ = ctx.last_kernel_times()
        work = ctx.me_work_per_frame(1)
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
    if rank != 0:
        return None
    allp = sorted(x for lst in gathered for x in lst)
    assert [g for g, _, _ in allp] == list(range(NGOP))
    # frame k of a GOP searches min(k, NREF) references
    me_px = work * sum(min(k, NREF) for k in range(1, IP)) * len(mine)
    return {
        "workload": "BASELINE configs[4]: 3840x2160, 64 GOPs x 8 frames, i=16, r=64 full search, 4 refs, QP 4",
        "n_gpus": world, "gops_per_rank": len(mine), "frames": NGOP * IP, "seconds": dt_max, "frames_per_s": NGOP * IP / dt_max,
        "rank0_device_ms": clip_ms, "rank0_kernel_ms": {k: v[0] for k, v in kt.items()},
        "rank0_search_tpx_per_s": me_px / (kt["me"][0] * 1e-3) / 1e12 if kt["me"][0] > 0 else None,
        "note": "rank0_kernel_ms from an extra pass with one lane group (kernels serialised on one stream); host buffers in and out",
        "container_bytes": sum(n for _, _, n in allp),
        "container_sha256_of_gop_hashes": hashlib.sha256("".join(h for _, h, _ in allp).encode()).hexdigest(),
        "round1_sha256_of_gop_hashes": "bc989c8d7986c95cc76772050c91c9403da8271b36059ac92abf2a1f3a68ce5f",
    }


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure(rank, world, local)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
