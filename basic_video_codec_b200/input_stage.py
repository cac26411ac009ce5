"""Input stage: YUV 4:2:0 -> Y-only planes and padding (reference assign1/ex2.py:14-46, common.py:13-32).

Host file helpers with the reference's names (read_y_component, save_y_frames_to_file, calculate_num_frames,
pad_frame) plus encode_yuv_file(), which skips the intermediate .y file: the luma planes go from the I420 file image
to HBM with one strided copy (bvc_clip_upload_i420) and are padded on the device.
"""
from __future__ import annotations

import os

import numpy as np

from ._lib import Context
from .encoder.encoder import pad_frame  # noqa: F401  (same contract as common.pad_frame)


def frame_bytes_i420(width, height):
    return width * height + 2 * (width // 2) * (height // 2)


def calculate_num_frames(file_path, width, height):
    """common.py:13-19: file size // bytes per I420 frame."""
    return os.path.getsize(file_path) // frame_bytes_i420(width, height)


def read_y_component(file_path, width, height, num_frames):
    """assign1/ex2.py:14-28: yield the Y plane of each I420 frame."""
    y_size, uv_size = width * height, (width // 2) * (height // 2)
    with open(file_path, "rb") as fh:
        for _ in range(num_frames):
            y = np.frombuffer(fh.read(y_size), dtype=np.uint8).reshape((height, width))
            fh.read(uv_size)
            fh.read(uv_size)
            yield y


def save_y_frames_to_file(params, frames_to_extract=None):
    """assign1/ex2.py:31-46: write <prefix>.y from <prefix>.yuv unless it already exists."""
    if getattr(params, "yuv_file", None) is None:
        params.yuv_file = os.path.splitext(params.y_only_file)[0] + ".yuv"
    n = frames_to_extract if frames_to_extract else calculate_num_frames(params.yuv_file, params.width, params.height)
    if os.path.exists(params.y_only_file):
        return
    with open(params.y_only_file, "wb") as out:
        for y in read_y_component(params.yuv_file, params.width, params.height, n):
            out.write(y.tobytes())


def encode_yuv_file(yuv_path, width, height, encoder_config, frames_to_process=None, device=0, max_lanes=None):
    """Encode the luma of an I420 file with the clip path (RCflag = 0).  Returns the container bytes
    (== encode_video's encoded.bin for the extracted, padded Y planes)."""
    ec = encoder_config
    if getattr(ec, "RCflag", 0):
        raise NotImplementedError("rate control (RCflag != 0) runs through encode_video")
    n = calculate_num_frames(yuv_path, width, height)
    if frames_to_process:
        n = min(n, frames_to_process)
    bs = ec.block_size
    W, H = width + (-width) % bs, height + (-height) % bs
    yuv = np.fromfile(yuv_path, dtype=np.uint8, count=n * frame_bytes_i420(width, height))
    ngop = (n + ec.I_Period - 1) // ec.I_Period
    with Context(W, H, bs, ec.search_range, ec.quantization_factor, ec.nRefFrames, ec.fastME, ec.fracMeEnabled, ec.I_Period,
                 device=device, max_lanes=max_lanes or min(ngop, 32)) as ctx:
        ctx.clip_upload_i420(yuv, width, height, n)
        out, ln = ctx.encode_clip_resident(n)
        return out[:ln].tobytes()
