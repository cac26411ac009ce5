"""bvc_decode_clip on the headline clip with pinned buffers: with the decoded planes returned, and kernels only.
Usage: python profiles/exp_decode_time.py"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import basic_video_codec_b200 as bvc
from basic_video_codec_b200._lib import _p
from tests import synth
W, H, BS, R, QP, IP, N = 1920, 1088, 16, 32, 4, 30, 600
frames = synth.moving_clip(11, H, W, N, step=6, clamp=96)
with bvc.Context(W, H, BS, R, QP, 1, False, False, IP, device=0, max_lanes=20) as ctx:
    out = torch.empty(N * W * H // 2, dtype=torch.uint8, pin_memory=True).numpy()
    ln = ctx.encode_clip_into(frames, out)
    data = out[:ln]
    dec = torch.empty((N, H, W), dtype=torch.uint8, pin_memory=True).numpy()
    n = C.c_int(0)
    for label, fo in (("with planes", dec), ("kernels only", None)):
        for it in range(3):
            t0 = time.perf_counter()
            rc = ctx._L.bvc_decode_clip(ctx._h, _p(data), data.size, N, _p(fo) if fo is not None else None, C.byref(n), None, None, None, None)
            dt = time.perf_counter() - t0
        print(label, rc, n.value, f"{dt * 1e3:.2f} ms")
