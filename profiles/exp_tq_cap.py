#!/usr/bin/env python3
"""P transform beside the other lane group's search: clip time against the CTA cap of tq_pframe_kernel (BVC_TQ_CTAS, 0 = one
CTA per work unit) and the number of lane groups.  Headline geometry, clip resident in HBM, stream hash checked.
Usage: python profiles/exp_tq_cap.py [lanes ...]"""
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
lanes_list = [int(x) for x in sys.argv[1:]] or [20, 3]
caps = [int(x) for x in os.environ.get("CAPS", "0,148,296,444,592,888").split(",")]
groups_list = [int(x) for x in os.environ.get("LANE_GROUPS", "2,3,4").split(",")]
base = synth.moving_clip(1080, H, W, IP * max(lanes_list), step=6, clamp=96, noise=2)
for lanes in lanes_list:
    n = lanes * IP
    frames = base[:n]
    out = np.empty(n * W * H // 2, np.uint8)
    ref = None
    for cap in caps:
        os.environ["BVC_TQ_CTAS"] = str(cap)
        with bvc.Context(W, H, BS, R, QP, 1, False, False, IP, device=0, max_lanes=lanes) as ctx:
            ctx.clip_upload(frames)
            for groups in groups_list:
                if groups > lanes:
                    continue
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
                print(json.dumps({"lanes": lanes, "tq_cta_cap": cap, "lane_groups": groups, "frames": n, "ms_per_clip": round(dt * 1e3, 3),
                                  "frames_per_s": round(n / dt, 1), "same_stream": h == ref}), flush=True)
