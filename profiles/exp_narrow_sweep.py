#!/usr/bin/env python3
"""Narrow-range search kernel (me_narrow.cu) on a GPU-filling frame: 1080p, i=16 / i=8, r = 1..7, integer and half-pel;
bvc_me_search on one frame (1 lane) and clip launches are both dominated by other things, so the kernel is timed here through
a 20-lane clip with one lane group (bvc_last_kernel_times, CUDA events around every search launch).
One JSON line per case: ms per launch, Tpx-absdiff/s, fraction of the VABSDIFF4 peak measured in the same process."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import basic_video_codec_b200 as bvc  # noqa: E402
from basic_video_codec_b200._lib import measure_peaks  # noqa: E402
from tests import synth  # noqa: E402

W, H, IP, LANES = 1920, 1088, 4, 20
cases = [(16, r, f) for r in (2, 3, 4, 5, 6, 7) for f in (False, True)] + [(8, r, f) for r in (1, 2, 3) for f in (False, True)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
    cases = [(b, r, bool(f)) for b, r, f in cases]
peak = measure_peaks(0)["px_absdiff_per_s"]
frames = synth.moving_clip(7, H, W, IP * LANES, step=2, clamp=32)
out = np.empty(frames.size, np.uint8)
for bs, r, frac in cases:
    with bvc.Context(W, H, bs, r, 4, 1, False, frac, IP, device=0, max_lanes=LANES) as ctx:
        ctx.set_lane_groups(1)
        ctx.clip_upload(frames)
        for _ in range(2):
            ctx.encode_clip_resident(frames.shape[0], out)
        acc_ms = acc_n = 0
        for _ in range(5):
            ctx.encode_clip_resident(frames.shape[0], out)
            kt, _ = ctx.last_kernel_times()
            acc_ms += kt["me"][0]
            acc_n += kt["me"][1]
        work = ctx.me_work_per_frame(1) * LANES
    ms = acc_ms / acc_n
    print(json.dumps({"bs": bs, "r": r, "half_pel": frac, "lanes": LANES, "ms_per_launch": ms, "tpx_per_s": work / (ms * 1e-3) / 1e12,
                      "frac_of_measured_peak": work / (ms * 1e-3) / peak, "peak_tpx_per_s": peak / 1e12}), flush=True)
