"""GOP-batched clip encoding (throughput path): the encode_video frame loop for RCflag = 0 with the
whole clip on the GPU and independent GOPs encoded in lock-step lanes."""
from __future__ import annotations

import numpy as np

from ._lib import Context


def encode_clip(frames: np.ndarray, encoder_config, device: int = 0, max_lanes: int | None = None,
                want_recon: bool = False):
    """Encode `frames` (n, H, W uint8, already padded to multiples of block_size) and return
    (container_bytes, recon or None).  container_bytes == the reference's encoded.bin for the same
    input and EncoderConfig (RCflag = 0)."""
    ec = encoder_config
    if getattr(ec, "RCflag", 0):
        raise NotImplementedError("rate control (RCflag != 0) is not part of the clip path")
    n, H, W = frames.shape
    ngop = (n + ec.I_Period - 1) // ec.I_Period
    lanes = max_lanes or min(ngop, 32)
    with Context(W, H, ec.block_size, ec.search_range, ec.quantization_factor, ec.nRefFrames, ec.fastME,
                 ec.fracMeEnabled, ec.I_Period, device=device, max_lanes=lanes) as ctx:
        return ctx.encode_clip(frames, want_recon=want_recon)
