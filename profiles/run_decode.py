#!/usr/bin/env python3
"""Decoder throughput on the headline clip (1920x1088, 600 frames, i=16, I_Period=30, QP=4): encode on the GPU,
decode the container on the GPU (bvc_decode_clip, host buffers in and out), check decode == the encoder's
reconstruction for all 600 frames (the full-size round-trip property), and time a CPU sample with the oracle.
Usage: python profiles/run_decode.py  -> one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import basic_video_codec_b200 as bvc  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from tests import synth  # noqa: E402

W, H, BS, R, QP, IP, N, LANES = 1920, 1088, 16, 32, 4, 30, 600, 20


def main():
    import torch
    frames = synth.moving_clip(1080, H, W, N, step=6, clamp=96, noise=2)
    out = np.empty(N * W * H // 2, np.uint8)
    recon = np.empty_like(frames)
    dec_t = torch.empty((N, H, W), dtype=torch.uint8, pin_memory=True)
    dec = dec_t.numpy()
    with bvc.Context(W, H, BS, R, QP, 1, False, False, IP, device=0, max_lanes=LANES) as ctx:
        ln = ctx.encode_clip_into(frames, out, recon)
        data_t = torch.empty(int(ln), dtype=torch.uint8, pin_memory=True)   # pinned host buffers in and out
        data = data_t.numpy()
        data[:] = out[:ln]
        l0 = ctx.launch_count()
        got = ctx.decode_clip(data, N, out=dec)       # warm-up (allocations)
        launches = ctx.launch_count() - l0
        same = bool(np.array_equal(got, recon))
        reps = 3
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.decode_clip(data, N, out=dec)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
    # CPU sample: the first GOP with the oracle decoder, one thread
    cfg = ob.make_config(W, H, BS, R, QP, nref=1, i_period=IP)
    from basic_video_codec_b200.sharding import split_container_by_gop
    gop0 = split_container_by_gop(data.tobytes(), [IP] * (N // IP))[0]
    t0 = time.perf_counter()
    d0 = ob.decode_clip(cfg, gop0, IP)
    dtc = time.perf_counter() - t0
    print(json.dumps({
        "workload": "decode of the headline clip: 1920x1088, 600 frames, i=16, I_Period=30, QP=4, %d GOP lanes" % LANES,
        "container_bytes": int(ln), "decode_ms": dt * 1e3, "decoded_frames_per_s": N / dt, "gpu_launches_per_clip": int(launches),
        "h2d_bytes": int(ln), "d2h_bytes": int(N * W * H), "roundtrip_equals_encoder_recon_all_frames": same,
        "cpu_oracle": {"frames": int(d0.shape[0]), "seconds": dtc, "frames_per_s": d0.shape[0] / dtc, "cores": 1,
                       "equals_gpu": bool(np.array_equal(d0, got[:IP]))},
    }), flush=True)


if __name__ == "__main__":
    main()
