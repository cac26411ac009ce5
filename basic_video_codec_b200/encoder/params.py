"""EncoderConfig: same constructor, attributes and validation as the reference (encoder/params.py:6-36)."""
import math


class EncoderConfig:
    def __init__(self, block_size, search_range, I_Period, quantization_factor, nRefFrames=1, fastME=False,
                 fracMeEnabled=False, RCflag=0, targetBR=0, resolution=(352, 288)):
        self.block_size = block_size
        self.search_range = search_range
        self.quantization_factor = quantization_factor
        self.I_Period = I_Period
        self.residual_approx_factor = 0
        self.nRefFrames = nRefFrames
        self.fastME = fastME
        self.fracMeEnabled = fracMeEnabled
        self.RCflag = RCflag
        self.rc_lookup_table = None
        self.targetBR = targetBR
        self.resolution = resolution
        self.frame_rate = 30
        self.validate()

    def validate(self):
        limit = math.log2(self.block_size) + 7
        if self.quantization_factor > limit:  # params.py:29-30
            raise ValueError(f" qp [{self.quantization_factor}] > {limit}")
        if self.RCflag and self.targetBR == 0:  # params.py:31-33
            raise ValueError("Target Bit Rate is 0 when Rate Control is On")
        if self.fastME:  # params.py:34-35: the search range is meaningless for FastME
            self.search_range = -1
        return self
