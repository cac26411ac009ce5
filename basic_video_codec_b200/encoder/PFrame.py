"""PFrame: inter frame with the reference's constructor and result attributes (encoder/PFrame.py:22-97)."""
from .Frame import Frame, context_for
from .PredictionMode import PredictionMode


class PFrame(Frame):
    def __init__(self, curr_frame=None, reference_frames=None, interpolated_reference_frames=None):
        super().__init__(curr_frame, reference_frames, interpolated_reference_frames)
        self.prediction_mode = PredictionMode.INTER_FRAME
        self.mv_field = {}
        self.avg_mae = None

    def encode_mc_q_dct(self, encoder_config):
        ec = encoder_config
        H, W = self.curr_frame.shape
        ctx = context_for(ec, W, H, self.device)
        # the half-pel planes are rebuilt on the GPU from the references; the reference's
        # interpolated_reference_frames deque is accepted for signature compatibility only
        r = self._encode_on(ctx, ec, list(self.reference_frames))
        self._store(r)
        self.residual_frame = r.resid_mc
        self.residual_wo_mc_frame = r.resid_nomc
        bs, bw = ec.block_size, W // ec.block_size
        fast = bool(ec.fastME)
        # raster order (= sorted by (y, x)); full search stores lists, FastME tuples (block_predictor.py:50-56,91)
        self.mv_field = {}
        for b in range(r.mv.shape[0]):
            key = ((b % bw) * bs, (b // bw) * bs)
            v = (int(r.mv[b, 0]), int(r.mv[b, 1]), int(r.mv[b, 2]))
            self.mv_field[key] = v if fast else list(v)
        return self
