"""Frame base class: the attribute protocol encoder/encoder.py relies on (reference encoder/Frame.py:22-48,
119-167).  The arithmetic lives in libbvc_b200.so; this class only carries results."""
from __future__ import annotations

import threading
from statistics import mean

import numpy as np

from .._lib import Context
from .PredictionMode import PredictionMode
from .RateControl.RateControl import (calculate_constant_row_bit_budget, calculate_proportional_row_bit_budget,
                                      find_rc_qp_for_row)


class BitString:
    """Read-only stand-in for the `bitarray` results the reference exposes: len() in bits, tobytes()
    zero-padded to whole bytes, truthiness, to01()."""
    __slots__ = ("_bytes", "_nbits")

    def __init__(self, data: bytes = b"", nbits: int = 0):
        self._bytes, self._nbits = bytes(data), int(nbits)

    def __len__(self):
        return self._nbits

    def __bool__(self):
        return self._nbits > 0

    def tobytes(self):
        return self._bytes

    def to01(self):
        return "".join(format(b, "08b") for b in self._bytes)[: self._nbits]

    def __eq__(self, other):
        return isinstance(other, BitString) and self._nbits == other._nbits and self._bytes == other._bytes


_ctx_cache: dict = {}
_ctx_lock = threading.Lock()


def context_for(ec, width, height, device=0, max_lanes=1) -> Context:
    """One cached GPU context per (geometry, parameters, device, lanes)."""
    key = (width, height, ec.block_size, ec.search_range, ec.quantization_factor, ec.nRefFrames, bool(ec.fastME),
           bool(ec.fracMeEnabled), ec.I_Period, device, max_lanes)
    with _ctx_lock:
        ctx = _ctx_cache.get(key)
        if ctx is None:
            ctx = Context(width, height, ec.block_size, ec.search_range, ec.quantization_factor, ec.nRefFrames,
                          ec.fastME, ec.fracMeEnabled, ec.I_Period, device=device, max_lanes=max_lanes)
            _ctx_cache[key] = ctx
        return ctx


class Frame:
    EOB_MARKER = 8190

    def __init__(self, curr_frame=None, reference_frames=None, interpolated_reference_frames=None):
        self.reference_frames = reference_frames
        self.interpolated_reference_frames = interpolated_reference_frames
        self.curr_frame = curr_frame
        self.prediction_mode = PredictionMode.INTER_FRAME
        self.entropy_encoded_prediction_data = None
        self.entropy_encoded_DCT_coffs = None
        self.residual_frame = None
        self.residual_wo_mc_frame = None
        self.quantized_dct_residual_frame = None
        self.reconstructed_frame = None
        self.avg_mae = None
        self.total_mae_comparisons = 0
        self.bit_budget = 0
        self.entropy_encoded_dct_length = 0
        self.entropy_encoded_prediction_data_length = 0
        self.rc_qp_per_row = []
        self.bits_per_row = []
        self.is_first_pass = True
        self.prev_pass_frame = None
        self.prev_frame = None
        self.index = 0
        self.scaling_factor = 1
        self.device = 0

    def encode_mc_q_dct(self, encoder_config):
        raise NotImplementedError(f"{type(self)} need to be overridden")

    # ---- rate control (reference Frame.py:155-188) ---------------------------------------------------
    def get_rc_qp(self, encoder_config, prev_frame_avg_qp, rc_qp, row_idx):
        ec = encoder_config
        frame_type = "I"  # Frame.py:169: `'I' if self.prediction_mode.INTRA_FRAME else 'P'` is always 'I'
        if ec.RCflag:
            if ec.RCflag == 1:
                budget = calculate_constant_row_bit_budget(self.bit_budget, row_idx, ec)
                rc_qp = find_rc_qp_for_row(budget, ec.rc_lookup_table, frame_type)
            if ec.RCflag > 1:
                if self.is_first_pass:
                    rc_qp = prev_frame_avg_qp
                else:
                    budget, _ = calculate_proportional_row_bit_budget(self, row_idx, ec)
                    rc_qp = find_rc_qp_for_row(budget, ec.rc_lookup_table, frame_type, scaling_factor=self.scaling_factor)
            self.rc_qp_per_row.append(rc_qp)
        return rc_qp

    def _prev_frame_avg_qp(self):
        """int(mean(prev_frame.rc_qp_per_row) - 0.1) + 1: a ceil with an offset of 0.1 (PFrame.py:49, IFrame.py:35).
        The reference crashes here on an empty list (second I frame with RCflag 0); the value is unused then."""
        q = getattr(self.prev_frame, "rc_qp_per_row", None)
        return int(mean(q) - 0.1) + 1 if q else 0

    def get_overage_ratios(self, encoder_config):
        """Frame size against the lookup's expected I / P frame size (Frame.py:155-163)."""
        ec = encoder_config
        if not self.is_first_pass:
            raise ValueError("why is overage being called in first pass?")
        bits = len(self.entropy_encoded_DCT_coffs) + len(self.entropy_encoded_prediction_data) + 8 * 6
        rows = ec.resolution[1] // ec.block_size
        row = ec.rc_lookup_table[ec.quantization_factor]
        return bits / (row["I"] * rows), bits / (row["P"] * rows)

    def _encode_on(self, ctx, ec, refs):
        """Run the frame on the GPU context with the reference's per-row QP protocol."""
        rows = self.curr_frame.shape[0] // ec.block_size
        rc_qp = ec.quantization_factor
        avg = self._prev_frame_avg_qp() if (ec.RCflag > 1 or refs is None) else 0
        if not ec.RCflag:
            return ctx.encode_iframe(self.curr_frame) if refs is None else ctx.encode_pframe(self.curr_frame, refs)
        if ec.RCflag == 1:
            # row k's QP depends on the bits rows < k consumed: encode row by row (bvc_frame_encode_row)
            ctx.frame_begin(self.curr_frame, refs)
            for row in range(rows):
                rc_qp = self.get_rc_qp(ec, avg, rc_qp, row)
                self.bit_budget -= ctx.frame_encode_row(row, rc_qp)
            r = ctx.frame_end()
            self.bit_budget += int(r.bits_per_row.sum())   # _store subtracts the row bits once more
            return r
        # RCflag 2 / 3: every row QP is known before the frame is coded
        qps = []
        for row in range(rows):
            rc_qp = self.get_rc_qp(ec, avg, rc_qp, row)
            qps.append(rc_qp)
        return ctx.encode_iframe(self.curr_frame, qps) if refs is None else ctx.encode_pframe(self.curr_frame, refs, qps)

    def _store(self, r):
        self.reconstructed_frame = r.recon
        self.quantized_dct_residual_frame = r.levels
        self.entropy_encoded_prediction_data = BitString(r.pred_bytes, r.pred_nbits)
        self.entropy_encoded_DCT_coffs = BitString(r.coef_bytes, r.coef_nbits)
        self.entropy_encoded_dct_length = r.coef_nbits
        self.entropy_encoded_prediction_data_length = r.pred_nbits
        self.bits_per_row = [int(b) for b in r.bits_per_row]
        self.bit_budget -= sum(self.bits_per_row)
        self.avg_mae = r.avg_mae
        self.total_mae_comparisons = r.mae_comparisons

    # ---- decoder side (reference Frame.py:81-110, PFrame.py:133-134,166-228, IFrame.py:85-114,132-166) ---------
    def _dec_ctx(self, params):
        ec = params.encoder_config
        bs = ec.block_size
        W, H = params.width + (-params.width) % bs, params.height + (-params.height) % bs
        return context_for(ec, W, H, self.device)

    def entropy_decode_prediction_data(self, enc, params):
        """Motion vectors / intra modes and the per-row QPs of one frame from its prediction payload."""
        ec = params.encoder_config
        ctx = self._dec_ctx(params)
        self._pred_payload = bytes(enc)
        pred, qps = ctx.decode_prediction_data(self.is_iframe(), self._pred_payload)
        self.rc_qp_per_row = [int(q) for q in qps]
        bs, bw = ec.block_size, ctx.W // ec.block_size
        if self.is_iframe():
            self.intra_modes = [int(m) for m in pred[:, 0]]
            return self.intra_modes
        self.mv_field = {((b % bw) * bs, (b // bw) * bs): (int(pred[b, 0]), int(pred[b, 1]), int(pred[b, 2])) for b in range(pred.shape[0])}
        return self.mv_field

    def entropy_decode_dct_coffs(self, params):
        """quantized_dct_residual_frame from self.entropy_encoded_DCT_coffs (bytes).  The GPU call decodes and
        rebuilds the frame in one go; decode_mc_q_dct then returns the rebuilt frame."""
        ctx = self._dec_ctx(params)
        refs = None if self.is_iframe() else [np.ascontiguousarray(r) for r in self.reference_frames]
        rec, lev, _, _ = ctx.decode_frame(self.is_iframe(), getattr(self, "_pred_payload", b""), bytes(self.entropy_encoded_DCT_coffs), refs)
        self.quantized_dct_residual_frame = lev
        self._decoded = rec
        return lev

    def decode_mc_q_dct(self, frame_shape, encoder_config):
        if getattr(self, "_decoded", None) is None:
            raise ValueError("call entropy_decode_prediction_data and entropy_decode_dct_coffs first")
        self.curr_frame = self._decoded
        return self._decoded

    def is_iframe(self):
        return self.prediction_mode == PredictionMode.INTRA_FRAME

    def is_pframe(self):
        return self.prediction_mode == PredictionMode.INTER_FRAME

    def get_quat_dct_coffs_extremes(self):
        q = self.quantized_dct_residual_frame
        if not isinstance(q, np.ndarray):
            raise TypeError("quantized_dct_residual_frame must be a numpy array")
        return [np.min(q), np.max(q)]

    def get_mv_extremes(self):
        if self.prediction_mode == PredictionMode.INTER_FRAME:
            arr = np.array(list(self.mv_field.values()))
            return [arr.min(axis=0).tolist(), arr.max(axis=0).tolist()]
        return [np.min(self.intra_modes), np.max(self.intra_modes)]

    def write_encoded_to_file(self, mv_fh, quant_dct_coff_fh, residual_yuv_fh, residual_wo_mc_yuv_fh, reconstructed_fh,
                              encoder_config):
        """Side files, same order and dtypes as the reference (Frame.py:119-130, file_io.py:65-74)."""
        residual_yuv_fh.write(self.residual_frame.tobytes())
        residual_wo_mc_yuv_fh.write(self.residual_wo_mc_frame.tobytes())
        quant_dct_coff_fh.write(self.quantized_dct_residual_frame.tobytes())
        reconstructed_fh.write(self.reconstructed_frame.tobytes())
        if self.prediction_mode == PredictionMode.INTER_FRAME:
            # sorted by (x, y) like file_io.write_mv_to_file; one write instead of one per block
            mvf = self.mv_field
            mv_fh.write("".join(f"{k[0]},{k[1]}:{mvf[k][0]},{mvf[k][1]}|" for k in sorted(mvf)))
        mv_fh.write("\n")
