#!/usr/bin/env python3
"""A clip ten times the benchmark's: 6000 frames of 1080p (12.5 GB of luma in pinned host memory; the 600-frame workload
clip repeated) through bvc_encode_clip.  Reports e2e frames/s, the device memory the context holds during the call
(cudaMemGetInfo before / after the first call) and checks the stream: every 600-frame slice must equal the 600-frame
stream (GOPs are independent).  Usage: python profiles/run_long_clip.py [repeats=10]"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import basic_video_codec_b200 as bvc  # noqa: E402
from tests import synth  # noqa: E402

W, H, BS, R, QP, IP, N = 1920, 1088, 16, 32, 4, 30, 600
rep = int(sys.argv[1]) if len(sys.argv) > 1 else 10
base = synth.moving_clip(1080, H, W, N, step=6, clamp=96, noise=2)
buf = torch.empty((rep * N, H, W), dtype=torch.uint8, pin_memory=True)
frames = buf.numpy()
for i in range(rep):
    frames[i * N:(i + 1) * N] = base
out_t = torch.empty(rep * N * W * H // 8, dtype=torch.uint8, pin_memory=True)
out = out_t.numpy()
torch.cuda.init()
free0, total = torch.cuda.mem_get_info()
with bvc.Context(W, H, BS, R, QP, 1, False, False, IP, device=0, max_lanes=20) as ctx:
    ln1 = ctx.encode_clip_into(frames[:N], out)           # the benchmark clip on its own
    sha1 = hashlib.sha256(out[:ln1].tobytes()).hexdigest()
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.encode_clip_into(frames[:N], out)
    fps1 = 3 * N / (time.perf_counter() - t0)
    ln = ctx.encode_clip_into(frames, out)                # warm-up of the long call (allocations)
    free1, _ = torch.cuda.mem_get_info()
    t0 = time.perf_counter()
    ln = ctx.encode_clip_into(frames, out)
    dt = time.perf_counter() - t0
    same = all(hashlib.sha256(out[i * ln1:(i + 1) * ln1].tobytes()).hexdigest() == sha1 for i in range(rep)) and ln == rep * ln1
print(json.dumps({"frames": rep * N, "input_bytes": int(frames.nbytes), "e2e_frames_per_s": rep * N / dt, "seconds": dt,
                  "e2e_frames_per_s_600_frame_clip": fps1, "device_bytes_held_by_context": int(free0 - free1),
                  "stream_bytes": int(ln), "every_600_frame_slice_equals_the_600_frame_stream": bool(same)}), flush=True)
assert same
