#!/usr/bin/env python3
"""The drop-in calls a user of the reference makes first -- encode_video(InputParameters) / decode_video(InputParameters),
frame objects, side files and all -- timed on BASELINE configs[0] (CIF 352x288, 10 frames, i=8, r=4, QP=3, I_Period=8; the
synthetic stand-in of tests/golden/cif_c1.npz, whose encoded.bin must come out byte-identical) and on a 30-frame 1080p
r=32 clip.  The reference's own encode_video needs about 1.7 s per CIF frame and about 440 s per 1080p P frame.
Usage: python profiles/run_dropin.py  -> one JSON line."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import basic_video_codec_b200 as bvc  # noqa: E402
from basic_video_codec_b200.encoder.encoder import encode_video, output_dir  # noqa: E402
from tests import golden_util as gu  # noqa: E402
from tests import synth  # noqa: E402


def run(frames, ec, td, name):
    n, H, W = frames.shape
    y = os.path.join(td, name + ".y")
    open(y, "wb").write(frames.tobytes())
    params = bvc.InputParameters(y, W, H, ec, frames_to_process=n)
    encode_video(params)                       # warm-up: context creation, module load
    t0 = time.perf_counter()
    encode_video(params)
    te = time.perf_counter() - t0
    bvc.decode_video(params)
    t0 = time.perf_counter()
    bvc.decode_video(params)
    tdec = time.perf_counter() - t0
    out = output_dir(params)
    enc = open(os.path.join(out, "encoded.bin"), "rb").read()
    same = open(os.path.join(out, "mc_decoded.yuv"), "rb").read() == open(os.path.join(out, "mc_reconstructed.yuv"), "rb").read()
    return {"frames": n, "encode_video_s": te, "encode_frames_per_s": n / te, "decode_video_s": tdec, "decode_frames_per_s": n / tdec,
            "decoded_equals_reconstructed": same}, enc


def main():
    res = {}
    with tempfile.TemporaryDirectory(prefix="bvc_dropin_") as td:
        g = gu.load("cif_c1")
        e = g["meta"]["enc"]
        n, H, W = g["frames"].shape
        ec = bvc.EncoderConfig(e["block"], e["search_range"], e["i_period"], e["qp"], nRefFrames=e.get("nref", 1), resolution=(W, H))
        r, enc = run(g["frames"], ec, td, "cif")
        r["encoded_bin_equals_reference_golden"] = enc == g["encoded"]
        res["config0_cif_i8_r4"] = r
        frames = synth.moving_clip(1080, 1088, 1920, 30, step=6, clamp=96, noise=2)
        ec = bvc.EncoderConfig(16, 32, 30, 4, nRefFrames=1, resolution=(1920, 1088))
        res["1080p_r32_30frames"] = run(frames, ec, td, "hd")[0]
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
