"""decode_video: drop-in for the reference's decoder.py:26-87 on the GPU decoder (bvc_decode_clip).

Same signature and files: reads <prefix>/<ident>/encoded.bin, writes mc_decoded.yuv next to it and -- like the
reference -- compares every decoded frame with mc_reconstructed.yuv (PSNR; identical frames give inf).
The whole container is decoded by one GPU call: the bit streams of all frames are tokenised in parallel, then the
GOPs (I-frame boundaries) are rebuilt several at a time.  A frame-by-frame variant with the reference's object
protocol (IFrame / PFrame .entropy_decode_prediction_data / .entropy_decode_dct_coffs / .decode_mc_q_dct) is
decode_video_framewise().  There is no CPU fallback.
"""
from __future__ import annotations

import logging
import os
from collections import deque

import numpy as np

from .encoder.encoder import _psnr, output_dir
from .encoder.Frame import context_for
from .encoder.IFrame import IFrame
from .encoder.PFrame import PFrame

logger = logging.getLogger("basic_video_codec_b200")


def _padded(params):
    bs = params.encoder_config.block_size
    return params.width + (-params.width) % bs, params.height + (-params.height) % bs


def decode_video(params, device: int = 0, max_lanes: int = 16):
    ec = params.encoder_config
    out = output_dir(params)
    W, H = _padded(params)
    data = open(os.path.join(out, "encoded.bin"), "rb").read()
    ctx = context_for(ec, W, H, device, max_lanes)       # cached: creating a context allocates the device pools
    frames = ctx.decode_clip(data, params.frames_to_process)
    rec_path = os.path.join(out, "mc_reconstructed.yuv")
    rec = np.fromfile(rec_path, dtype=np.uint8) if os.path.exists(rec_path) else None
    with open(os.path.join(out, "mc_decoded.yuv"), "wb") as fh:
        for i, fr in enumerate(frames):
            if rec is not None and rec.size >= (i + 1) * W * H:
                logger.info("%2d: psnr [%6.2f]", i + 1, _psnr(fr, rec[i * W * H:(i + 1) * W * H].reshape(H, W)))
            fh.write(fr.tobytes())
    logger.info("End decoding")


def decode_video_framewise(params, device: int = 0):
    """The reference's loop, frame object by frame object (decoder.py:44-85)."""
    ec = params.encoder_config
    out = output_dir(params)
    W, H = _padded(params)
    reference_frames = deque(maxlen=ec.nRefFrames)
    reference_frames.append(np.full((H, W), 128, dtype=np.uint8))
    with open(os.path.join(out, "encoded.bin"), "rb") as enc_fh, open(os.path.join(out, "mc_decoded.yuv"), "wb") as dec_fh:
        idx = 0
        while True:
            idx += 1
            t = enc_fh.read(1)
            if idx > params.frames_to_process or not t:
                break
            if t[0] == 1:
                frame = IFrame()
                reference_frames.clear()
            else:
                frame = PFrame(reference_frames=reference_frames, interpolated_reference_frames=None)
            frame.device = device
            frame.entropy_decode_prediction_data(enc_fh.read(int.from_bytes(enc_fh.read(2), "big")), params)
            frame.entropy_encoded_DCT_coffs = enc_fh.read(int.from_bytes(enc_fh.read(3), "big"))
            frame.entropy_decode_dct_coffs(params)
            decoded = frame.decode_mc_q_dct((H, W), ec)
            dec_fh.write(decoded.tobytes())
            reference_frames.append(decoded)
