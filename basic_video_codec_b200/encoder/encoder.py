"""encode_video: the reference's entry point (encoder/encoder.py:28-171) re-driven over the GPU frames.

Same signature, same output directory scheme and files (file_io.py:20-26): encoded.bin,
mc_reconstructed.yuv, mc_quant_dct_coff.bin, residuals_w_mc.yuv, residuals_wo_mc.yuv, mv.txt,
metrics.csv.  Differences, all deliberate: no rate-control lookup CSV is required when RCflag = 0
(reference quirk Q14), nothing is appended to a results.csv inside the package, and the half-pel
planes live on the GPU.  Rate control (RCflag 1: per-row feedback; 2/3: two passes with scene-change
detection, encoder.py:85-98,188-201) is driven from here exactly like the reference does; its lookup
tables are only needed when RCflag != 0.
"""
from __future__ import annotations

import csv
import os
import time
from collections import deque

import numpy as np

from .Frame import Frame
from .IFrame import IFrame
from .PFrame import PFrame
from .RateControl.RateControl import bit_budget_per_frame
from .RateControl.lookup import get_combined_lookup_table, rc_lookup_file_path


def pad_frame(frame, block_size, pad_value=128):
    """common.pad_frame (common.py:22-32): pad bottom/right with 128 to multiples of block_size."""
    h, w = frame.shape
    ph, pw = (-h) % block_size, (-w) % block_size
    if ph or pw:
        out = np.full((h + ph, w + pw), pad_value, dtype=np.uint8)
        out[:h, :w] = frame
        return out
    return frame


def _psnr(a, b):
    """skimage.metrics.peak_signal_noise_ratio for uint8 planes (encoder.py:9,123).  The sum of squared differences is
    taken in integers: every partial sum of the reference's float64 mean is an integer below 2^53, so the value is the
    same to the last bit, without three float64 temporaries per frame."""
    d = a.astype(np.int16) - b.astype(np.int16)
    sse = int(np.sum(np.multiply(d, d, dtype=np.int32), dtype=np.int64))
    if sse == 0:
        return float("inf")
    mse = np.float64(sse) / np.float64(a.size)
    return 10 * np.log10(255.0 ** 2 / mse)


def output_dir(params):
    ec = params.encoder_config
    fme = ".0" if ec.fracMeEnabled else ""
    ident = f"{ec.block_size}_{ec.search_range}{fme}_{ec.quantization_factor}_{ec.I_Period}_{ec.nRefFrames}_{ec.RCflag}_{ec.targetBR}"
    return os.path.join(os.path.splitext(params.y_only_file)[0], ident)


def encode_video(params, device: int = 0):
    ec = params.encoder_config
    if ec.RCflag:
        ec.rc_lookup_table = get_combined_lookup_table(rc_lookup_file_path(ec, "I"), rc_lookup_file_path(ec, "P"))
    scene_change_threshold = 1.3
    out = output_dir(params)
    os.makedirs(out, exist_ok=True)
    W, H, bs = params.width, params.height, ec.block_size
    reference_frames = deque(maxlen=ec.nRefFrames)
    reference_frames.append(np.full((H, W), 128, dtype=np.uint8))
    interpolated_reference_frames = deque(maxlen=ec.nRefFrames)  # kept for signature compatibility
    t_start = time.time()
    with open(params.y_only_file, "rb") as f_in, \
            open(os.path.join(out, "mv.txt"), "wt") as mv_fh, \
            open(os.path.join(out, "mc_quant_dct_coff.bin"), "wb") as coef_fh, \
            open(os.path.join(out, "residuals_w_mc.yuv"), "wb") as res_fh, \
            open(os.path.join(out, "residuals_wo_mc.yuv"), "wb") as res0_fh, \
            open(os.path.join(out, "mc_reconstructed.yuv"), "wb") as rec_fh, \
            open(os.path.join(out, "encoded.bin"), "wb") as enc_fh, \
            open(os.path.join(out, "metrics.csv"), "wt", newline="") as met_fh:
        met = csv.writer(met_fh)
        met.writerow(["idx", "I-Frame", "avg_MAE", "mae_comps", "PSNR", "frame_bytes", "file_bits", "enc_time", "elapsed_time"])
        prev = Frame()
        prev.rc_qp_per_row = [ec.quantization_factor]  # arbitrary seed for the first frame (encoder.py:72-73)
        idx = 0
        while True:
            t0 = time.time()
            start = enc_fh.tell()
            idx += 1
            raw = f_in.read(W * H)
            if not raw or idx > params.frames_to_process:
                break
            cur = pad_frame(np.frombuffer(raw, dtype=np.uint8).reshape(H, W), bs)

            def make(intra, first_pass, prev_pass=None):
                if intra:  # encoder.py:174-178,189-192
                    reference_frames.clear()
                    interpolated_reference_frames.clear()
                    fr = IFrame(cur)
                else:
                    fr = PFrame(cur, reference_frames, interpolated_reference_frames)
                fr.is_first_pass, fr.prev_frame, fr.index, fr.device = first_pass, prev, idx, device
                fr.bit_budget = bit_budget_per_frame(ec) if ec.RCflag else 0
                fr.prev_pass_frame = prev_pass
                return fr

            frame = make((idx - 1) % ec.I_Period == 0, True)
            frame.encode_mc_q_dct(ec)
            if ec.RCflag > 1:  # second pass, encoder.py:91-98
                first = frame
                overage = first.get_overage_ratios(ec)
                scene_change = False
                if first.is_pframe() and overage[1] > scene_change_threshold:
                    first.scaling_factor = (1 - overage[1]) * 0.95  # set on the first-pass object, as the reference does
                    scene_change = True
                frame = make(scene_change or first.is_iframe(), False, first)
                frame.encode_mc_q_dct(ec)
            enc_time = time.time() - t0
            # container, encoder.py:104-121
            pb = (len(frame.entropy_encoded_prediction_data) + 7) // 8
            cb = (len(frame.entropy_encoded_DCT_coffs) + 7) // 8
            enc_fh.write(frame.prediction_mode.value.to_bytes(1, "big"))
            enc_fh.write(pb.to_bytes(2, "big"))
            enc_fh.write(frame.entropy_encoded_prediction_data.tobytes())
            enc_fh.write(cb.to_bytes(3, "big"))
            enc_fh.write(frame.entropy_encoded_DCT_coffs.tobytes())
            psnr = _psnr(frame.curr_frame, frame.reconstructed_frame)
            met.writerow([idx, frame.prediction_mode.value, frame.avg_mae, frame.total_mae_comparisons, psnr,
                          enc_fh.tell() - start, enc_fh.tell() * 8, enc_time, time.time() - t_start])
            frame.write_encoded_to_file(mv_fh, coef_fh, res_fh, res0_fh, rec_fh, ec)
            reference_frames.append(frame.reconstructed_frame)
            prev = frame
    return
