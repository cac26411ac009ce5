"""PFrame: inter frame with the reference's constructor and result attributes (encoder/PFrame.py:22-97)."""
from .Frame import Frame, context_for
from .PredictionMode import PredictionMode


_KEYS = {}


def _block_keys(W, H, bs):
    """[(x, y)] of every block in raster order, cached per geometry."""
    k = _KEYS.get((W, H, bs))
    if k is None:
        k = _KEYS[(W, H, bs)] = [(x, y) for y in range(0, H, bs) for x in range(0, W, bs)]
    return k


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
        # raster order (= sorted by (y, x)); full search stores lists, FastME tuples (block_predictor.py:50-56,91)
        keys = _block_keys(W, H, bs)
        vals = r.mv.tolist()
        self.mv_field = dict(zip(keys, map(tuple, vals))) if ec.fastME else dict(zip(keys, vals))
        return self
