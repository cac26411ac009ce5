"""The reference's own encoder / decoder test intents (tests/test_encoder.py:37-203, tests/test_decoder.py:37-78),
re-plumbed onto the current frame protocol (reference windows are deques; the reference's tests still call the
pre-deque constructor and no longer import) and run on the GPU layer: a marker moved by (tx, ty) must be found with
that motion vector, and decoding the frame's own streams must give back its reconstruction."""
from collections import deque

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _marked(f_size, x0, y0, size, fill):
    f = np.zeros((f_size, f_size), dtype=np.uint8)
    f[y0:y0 + size, x0:x0 + size] = fill
    return f


def _encode_p(cur, prev, ec):
    from basic_video_codec_b200.encoder.PFrame import PFrame
    fr = PFrame(cur, deque([prev], maxlen=1), deque(maxlen=1))
    return fr.encode_mc_q_dct(ec)


def test_encode_frame_right_down_motion():
    """tests/test_encoder.py:77-136: content moved by (-1,-1) => mv (1,1) for the block holding the marker."""
    from basic_video_codec_b200 import EncoderConfig
    bs, r, nb = 8, 3, 3
    f = bs * nb
    ec = EncoderConfig(bs, r, I_Period=8, quantization_factor=0, resolution=(f, f))
    for bx in range(nb - 1):
        for by in range(nb - 1):
            prev = _marked(f, bs * bx + 1, bs * by + 2, 1, 69)
            cur = np.roll(np.roll(prev, -1, axis=1), -1, axis=0)
            enc = _encode_p(cur, prev, ec)
            mv = enc.mv_field[(bs * bx, bs * by)]
            assert (mv[0], mv[1]) == (1, 1), (bx, by, mv)
            assert enc.avg_mae <= 5


def test_encode_frame_left_up_motion():
    """tests/test_encoder.py:138-203: marker with an added residual, moved by (+1,+1) => mv (-1,-1)."""
    from basic_video_codec_b200 import EncoderConfig
    bs, r, nb = 8, 3, 4
    f = bs * nb
    ec = EncoderConfig(bs, r, I_Period=8, quantization_factor=0, resolution=(f, f))
    for bx in range(1, nb):
        for by in range(1, nb):
            prev = _marked(f, bs * bx + 1, bs * by + 1, 2, 42)
            cur = np.roll(np.roll(_marked(f, bs * bx + 1, bs * by + 1, 2, 42 + 7), 1, axis=1), 1, axis=0)
            enc = _encode_p(cur, prev, ec)
            mv = enc.mv_field[(bs * bx, bs * by)]
            assert (mv[0], mv[1]) == (-1, -1), (bx, by, mv)
            assert enc.avg_mae <= 5


def test_decode_frame_right_down_motion():
    """tests/test_decoder.py:37-78: decode(levels, mv) of an encoded P frame equals its reconstruction (the reference
    allows +-2; both sides use the same defined transform here, so it is exact)."""
    from basic_video_codec_b200 import EncoderConfig, InputParameters
    from basic_video_codec_b200.encoder.PFrame import PFrame
    bs, r, nb, qp = 4, 3, 4, 8
    f = bs * nb
    ec = EncoderConfig(bs, r, 1, qp, resolution=(f, f))
    params = InputParameters("unused.y", f, f, ec, frames_to_process=1)
    for bx in range(nb):
        for by in range(nb):
            prev = _marked(f, bs * bx + 1, bs * by + 1, 2, 99)
            cur = np.roll(np.roll(prev, 1, axis=1), 2, axis=0)
            enc = _encode_p(cur, prev, ec)
            dec = PFrame(reference_frames=deque([prev], maxlen=1), interpolated_reference_frames=None)
            mvs = dec.entropy_decode_prediction_data(enc.entropy_encoded_prediction_data.tobytes(), params)
            assert {k: tuple(v) for k, v in enc.mv_field.items()} == mvs
            dec.entropy_encoded_DCT_coffs = enc.entropy_encoded_DCT_coffs.tobytes()
            assert np.array_equal(dec.entropy_decode_dct_coffs(params), enc.quantized_dct_residual_frame)
            assert np.array_equal(dec.decode_mc_q_dct((f, f), ec), enc.reconstructed_frame)
