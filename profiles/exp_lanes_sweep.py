#!/usr/bin/env python3
"""Headline geometry (1080p, i=16, r=32) with few GOP lanes per GPU -- what a rank sees when the 20 GOPs of BASELINE
configs[3] are sharded over 8 GPUs (2-3 lanes).  Clip resident in HBM; frames/s and search-launch time per lane count,
with and without the tail split of the search grid (BVC_TAIL_SPLIT).  One JSON line per (lanes, tail_split).
Usage: python profiles/exp_lanes_sweep.py [lanes ...]"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import basic_video_codec_b200 as bvc  # noqa: E402
from tests import synth  # noqa: E402

W, H, BS, R, QP, IP = 1920, 1088, 16, 32, 4, 30
lanes_list = [int(x) for x in sys.argv[1:]] or [1, 2, 3, 5, 10, 20]
base = synth.moving_clip(1080, H, W, IP * max(lanes_list), step=6, clamp=96, noise=2)
for lanes in lanes_list:
    n = lanes * IP
    frames = base[:n]
    out = np.empty(n * W * H // 2, np.uint8)
    ref = None
    for split in (0, 1):
        os.environ["BVC_TAIL_SPLIT"] = str(split)
        with bvc.Context(W, H, BS, R, QP, 1, False, False, IP, device=0, max_lanes=lanes) as ctx:
            ctx.clip_upload(frames)
            for groups in ((1, 2) if lanes > 1 else (1,)):
                ctx.set_lane_groups(groups)
                for _ in range(2):
                    ctx.encode_clip_resident(n, out)
                reps = max(3, 60 // lanes)
                t0 = time.perf_counter()
                for _ in range(reps):
                    _, ln = ctx.encode_clip_resident(n, out)
                dt = (time.perf_counter() - t0) / reps
                h = hashlib.sha256(out[:ln].tobytes()).hexdigest()[:16]
                ref = ref or h
                kt, clip_ms = ctx.last_kernel_times()
                work = ctx.me_work_per_frame(1) * lanes
                me_ms = kt["me"][0] / max(1, kt["me"][1])
                print(json.dumps({"lanes": lanes, "tail_split": split, "lane_groups": groups, "frames": n, "ms_per_clip": dt * 1e3,
                                  "frames_per_s": n / dt, "me_launch_ms": me_ms if groups == 1 else None,
                                  "me_tpx_per_s": (work / (me_ms * 1e-3) / 1e12) if (groups == 1 and me_ms > 0) else None,
                                  "same_stream": h == ref}), flush=True)
