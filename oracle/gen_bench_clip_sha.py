#!/usr/bin/env python3
"""The CPU oracle's encoding of bench.py's whole workload clip (BASELINE configs[3]: 1920x1088, 600 frames, i=16, r=32,
I_Period=30, QP 4, seed 1080) -> tests/golden/bench_clip_oracle.json: sha256 and length of the serial stream and of every
GOP fragment.  bench.py compares the GPU stream's sha with it on every run (so the headline number is a number for a
bit-exact stream), tests/test_gpu_fullsize.py checks whole 30-frame GOPs through 20 lanes / 2 lane groups against the
per-GOP entries.  ~4 minutes on 8 cores.  TEST INFRASTRUCTURE ONLY.
Usage: python -m oracle.gen_bench_clip_sha"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402
from tests import synth  # noqa: E402
from basic_video_codec_b200.sharding import split_container_by_gop  # noqa: E402

W, H, BS, R, QP, IP, N = 1920, 1088, 16, 32, 4, 30, 600

if __name__ == "__main__":
    t0 = time.time()
    frames = synth.moving_clip(1080, H, W, N, step=6, clamp=96, noise=2)
    cfg = ob.make_config(W, H, BS, R, QP, nref=1, i_period=IP)
    data, _ = ob.encode_clip(cfg, frames, nthreads=os.cpu_count() or 1, want_recon=False)
    parts = split_container_by_gop(data, [IP] * (N // IP))
    out = {"workload": "synth.moving_clip(1080, 1088, 1920, 600, step=6, clamp=96, noise=2); i=16 r=32 qp=4 I_Period=30 nRef=1",
           "bytes": len(data), "sha256": hashlib.sha256(data).hexdigest(),
           "gops": [{"bytes": len(p), "sha256": hashlib.sha256(p).hexdigest()} for p in parts],
           "seconds": round(time.time() - t0, 1), "threads": os.cpu_count()}
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "bench_clip_oracle.json"), "w"), indent=1)
    print(out["bytes"], out["sha256"], out["seconds"])
