"""Frame base class: the attribute protocol encoder/encoder.py relies on (reference encoder/Frame.py:22-48,
119-167).  The arithmetic lives in libbvc_b200.so; this class only carries results."""
from __future__ import annotations

import threading

import numpy as np

from .._lib import Context
from .PredictionMode import PredictionMode


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


def context_for(ec, width, height, device=0) -> Context:
    """One cached GPU context per (geometry, parameters, device)."""
    key = (width, height, ec.block_size, ec.search_range, ec.quantization_factor, ec.nRefFrames, bool(ec.fastME),
           bool(ec.fracMeEnabled), ec.I_Period, device)
    with _ctx_lock:
        ctx = _ctx_cache.get(key)
        if ctx is None:
            ctx = Context(width, height, ec.block_size, ec.search_range, ec.quantization_factor, ec.nRefFrames,
                          ec.fastME, ec.fracMeEnabled, ec.I_Period, device=device, max_lanes=1)
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

    def _row_qps(self, ec):
        """Per-row QPs.  Rate control (RCflag != 0, Frame.py:168-188) feeds row bit counts back into the
        next row's QP; that loop is host logic outside this hot path."""
        if getattr(ec, "RCflag", 0):
            raise NotImplementedError("RCflag != 0: rate control is not part of the B200 hot path yet")
        return None

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
            for k in sorted(self.mv_field.keys()):
                mv_fh.write(f"{k[0]},{k[1]}:{self.mv_field[k][0]},{self.mv_field[k][1]}|")
        mv_fh.write("\n")
