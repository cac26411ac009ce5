#!/usr/bin/env python3
"""BASELINE configs[2] shape with rate control: CIF, i=16, r=4, half-pel, RCflag=1, 300 frames (I_Period 21 as in
assign3/Ex1.py, 2.4 Mbit/s, the reference's lookup tables) -- frames/s
  before: encode_video's frame loop, one launch + sync + read-back per block row (bvc_frame_encode_row)
  after : the clip call with the per-row feedback chained on the device (bvc_set_rate_control), GOP lanes in lock step
and a check that both produce the same stream.  One JSON line."""
import hashlib
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import basic_video_codec_b200 as bvc  # noqa: E402
from basic_video_codec_b200.encoder import encoder as enc_mod  # noqa: E402
from tests import synth  # noqa: E402

W, H, BS, R, QP, IP, N, BR = 352, 288, 16, 4, 4, 21, 294, 2_400_000
frames = synth.moving_clip(303, H, W, N, step=2, clamp=16)
ec = bvc.EncoderConfig(BS, R, IP, QP, nRefFrames=1, fracMeEnabled=True, RCflag=1, targetBR=BR, resolution=(W, H))

with tempfile.TemporaryDirectory() as td:
    yfile = os.path.join(td, "clip.y")
    open(yfile, "wb").write(frames.tobytes())
    params = bvc.InputParameters(yfile, W, H, ec, frames_to_process=N)
    enc_mod.encode_video(params)          # warm-up (context creation)
    t0 = time.perf_counter()
    enc_mod.encode_video(params)
    dt_loop = time.perf_counter() - t0
    ref = open(os.path.join(enc_mod.output_dir(params), "encoded.bin"), "rb").read()

from basic_video_codec_b200.clip import configure_rate_control  # noqa: E402
lanes = N // IP
out = np.empty(N * W * H, np.uint8)
with bvc.Context(W, H, BS, R, QP, 1, False, True, IP, device=0, max_lanes=lanes) as ctx:
    configure_rate_control(ctx, ec)
    ctx.encode_clip_into(frames, out)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ln = ctx.encode_clip_into(frames, out)
    dt_clip = (time.perf_counter() - t0) / reps
    launches = ctx.launch_count()
    ctx.set_rate_control(0)
    ctx.encode_clip_into(frames, out[: N * W * H // 2])
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.encode_clip_into(frames, out[: N * W * H // 2])
    dt_norc = (time.perf_counter() - t0) / reps
    configure_rate_control(ctx, ec)
    ln = ctx.encode_clip_into(frames, out)
same = out[:ln].tobytes() == ref
print(json.dumps({"workload": "CIF 352x288 i=16 r=4 half-pel RCflag=1 2.4 Mbit/s I_Period=21, 294 frames", "lanes": lanes,
                  "frame_loop_frames_per_s": N / dt_loop, "frame_loop_ms": dt_loop * 1e3,
                  "clip_call_frames_per_s": N / dt_clip, "clip_call_ms": dt_clip * 1e3,
                  "clip_call_without_rc_frames_per_s": N / dt_norc, "same_stream": same,
                  "stream_sha256": hashlib.sha256(ref).hexdigest(), "stream_bytes": len(ref)}), flush=True)
assert same
