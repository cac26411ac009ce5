"""IFrame: intra frame with the reference's constructor and result attributes (encoder/IFrame.py:16-83)."""
import numpy as np

from .Frame import Frame, context_for
from .PredictionMode import PredictionMode


class IFrame(Frame):
    def __init__(self, curr_frame=None):
        super().__init__(curr_frame)
        self.prediction_mode = PredictionMode.INTRA_FRAME
        self.intra_modes = None

    def encode_mc_q_dct(self, encoder_config):
        ec = encoder_config
        H, W = self.curr_frame.shape
        ctx = context_for(ec, W, H, self.device)
        r = self._encode_on(ctx, ec, None)
        self._store(r)
        self.intra_modes = [int(m) for m in r.modes]
        # IFrame.py:30,57-58,81-83: one uint8 plane serves as both residual planes
        self.residual_frame = r.resid_mc.view(np.uint8)
        self.residual_wo_mc_frame = self.residual_frame
        return self
