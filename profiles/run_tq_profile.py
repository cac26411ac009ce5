#!/usr/bin/env python3
"""Small driver for `ncu` captures of the P-frame transform kernel: 1080p, i=16, r=32, 10 GOP lanes of I + 5 P frames, one
lane group, clip resident.  Usage (on the GPU box):
  ncu --set full --clock-control none --import-source on -k regex:tq_pframe -s 12 -c 1 -o gpurun_out/NAME python profiles/run_tq_profile.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import basic_video_codec_b200 as bvc  # noqa: E402
from tests import synth  # noqa: E402

W, H, BS, R, QP, IP, LANES = 1920, 1088, 16, int(os.environ.get("TQP_R", "32")), 4, 6, 10
n = LANES * IP
frames = synth.moving_clip(1080, H, W, n, step=6, clamp=96, noise=2)
out = np.empty(n * W * H // 2, np.uint8)
with bvc.Context(W, H, BS, R, QP, 1, False, False, IP, device=0, max_lanes=LANES) as ctx:
    ctx.clip_upload(frames)
    ctx.set_lane_groups(1)
    for _ in range(3):
        _, ln = ctx.encode_clip_resident(n, out)
    kt, clip_ms = ctx.last_kernel_times()
    print({k: (v[0] / max(1, v[1])) for k, v in kt.items()}, clip_ms, ln)
