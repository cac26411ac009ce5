"""GOP-batched clip encoding (throughput path): the encode_video frame loop for RCflag = 0 with the
whole clip on the GPU and independent GOPs encoded in lock-step lanes."""
from __future__ import annotations

import numpy as np

from ._lib import Context


def configure_rate_control(ctx: Context, ec) -> None:
    """RCflag 1 of an EncoderConfig onto a context's clip path (bvc_set_rate_control): frame budget targetBR / frame_rate
    (encoder.py:181-185), table = ec.rc_lookup_table or the lookup CSV for ec.resolution / block size."""
    rcflag = getattr(ec, "RCflag", 0)
    if rcflag not in (0, 1):
        raise NotImplementedError("RCflag 2 / 3 (two passes, scene changes) couple consecutive GOPs: use encode_video")
    if not rcflag:
        return
    table = ec.rc_lookup_table
    if table is None:
        from .encoder.RateControl.lookup import get_combined_lookup_table, rc_lookup_file_path
        table = get_combined_lookup_table(rc_lookup_file_path(ec, "I"), rc_lookup_file_path(ec, "P"))
    from .encoder.RateControl.RateControl import bit_budget_per_frame
    ctx.set_rate_control(1, bit_budget_per_frame(ec), table)


def encode_clip(frames: np.ndarray, encoder_config, device: int = 0, max_lanes: int | None = None,
                want_recon: bool = False):
    """Encode `frames` (n, H, W uint8, already padded to multiples of block_size) and return
    (container_bytes, recon or None).  container_bytes == encoded.bin of the reference run with the defined fp64 DCT
     for the same input and EncoderConfig, RCflag 0 or 1.  RCflag 1 needs ec.rc_lookup_table (the
    reference's get_combined_lookup_table) or a lookup CSV for ec.resolution / block size; RCflag 2 / 3 couple GOPs and
    run through encode_video only."""
    ec = encoder_config
    rcflag = getattr(ec, "RCflag", 0)
    if rcflag not in (0, 1):
        raise NotImplementedError("RCflag 2 / 3 (two passes, scene changes) couple consecutive GOPs: use encode_video")
    n, H, W = frames.shape
    ngop = (n + ec.I_Period - 1) // ec.I_Period
    lanes = max_lanes or min(ngop, 32)
    with Context(W, H, ec.block_size, ec.search_range, ec.quantization_factor, ec.nRefFrames, ec.fastME,
                 ec.fracMeEnabled, ec.I_Period, device=device, max_lanes=lanes) as ctx:
        configure_rate_control(ctx, ec)
        return ctx.encode_clip(frames, want_recon=want_recon)
