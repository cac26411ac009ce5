"""GPU decoder (bvc_decode_clip / bvc_decode_frame) against the decoder oracle, the committed goldens (whose decoded
output was pinned with the reference's own decode_video, tests/golden/decode_ref.json) and the encoder's reconstruction."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import golden_util as gu
from tests import synth

pytestmark = pytest.mark.gpu


def _ob():
    from oracle import bindings as ob
    return ob


def _ctx(W, H, e, lanes=1):
    import basic_video_codec_b200 as bvc
    return bvc.Context(W, H, e["block"], e["search_range"], e["qp"], e.get("nref", 1), e.get("fastme", False), e.get("frac", False),
                       e["i_period"], device=0, max_lanes=lanes)


@pytest.mark.parametrize("name", gu.names())
def test_decode_golden_streams(name):
    """Streams written by the Python reference (RCflag 0..3, 1-4 refs, half-pel, FastME, i = 4/8/16)."""
    ob = _ob()
    ref = json.load(open(os.path.join(gu.GOLD, "decode_ref.json")))[name]
    g = gu.load(name)
    e = g["meta"]["enc"]
    n, H, W = g["frames"].shape
    cfg = ob.make_config(W, H, e["block"], e["search_range"], e["qp"], nref=e.get("nref", 1), fastme=e.get("fastme", False),
                         frac=e.get("frac", False), i_period=e["i_period"])
    want = ob.decode_clip(cfg, g["encoded"], n, details=True)
    for lanes in (1, 3):
        with _ctx(W, H, e, lanes) as ctx:
            got = ctx.decode_clip(g["encoded"], n + 2, details=True)
            assert got[0].shape[0] == n == ref["frames"]
            assert hashlib.sha256(got[0].tobytes()).hexdigest() == ref["decoded_sha256"], f"lanes={lanes}"
            for a, b, what in zip(got, want, ("frames", "levels", "pred", "qp_rows", "kinds")):
                assert np.array_equal(a, b), what
            if "recon" in g:
                assert np.array_equal(got[0], g["recon"])
            # frames_to_process cuts the loop (decoder.py:49)
            assert np.array_equal(ctx.decode_clip(g["encoded"], 2), want[0][:2])


def test_decode_roundtrip_many_gops():
    """encode on the GPU -> decode on the GPU == the encoder's reconstruction; GOP lanes, a short last GOP, 2 refs."""
    import basic_video_codec_b200 as bvc
    ob = _ob()
    H, W, bs, r, qp, ip, n = 64, 96, 16, 8, 2, 4, 23
    frames = synth.moving_clip(31, H, W, n, step=4, clamp=24)
    for frac in (False, True):
        with bvc.Context(W, H, bs, r, qp, 2, False, frac, ip, device=0, max_lanes=4) as ctx:
            data, recon = ctx.encode_clip(frames, want_recon=True)
            dec = ctx.decode_clip(data, n)
            assert np.array_equal(dec, recon)
        cfg = ob.make_config(W, H, bs, r, qp, nref=2, frac=frac, i_period=ip)
        assert np.array_equal(ob.decode_clip(cfg, data, n), recon)


def test_decode_stream_starting_with_p_frames_uses_the_128_window():
    """decoder.py:34-38: before the first I frame the reference window holds one plane of 128s."""
    import basic_video_codec_b200 as bvc
    ob = _ob()
    H, W, bs, r, qp, ip, n = 48, 64, 8, 4, 3, 3, 7
    frames = synth.moving_clip(5, H, W, n, step=2, clamp=8)
    for frac, nref in ((False, 1), (True, 2)):
        with bvc.Context(W, H, bs, r, qp, nref, False, frac, ip, device=0, max_lanes=2) as ctx:
            data, _ = ctx.encode_clip(frames)
            recs = gu.split_container(data)
            # drop the first I frame: the stream now opens with two P frames whose MVs point into the 128 plane
            cut = b"".join(bytes([m]) + len(p).to_bytes(2, "big") + p + len(c).to_bytes(3, "big") + c for m, p, c in recs[1:])
            cfg = ob.make_config(W, H, bs, r, qp, nref=nref, frac=frac, i_period=ip)
            want = ob.decode_clip(cfg, cut, n, details=True)
            got = ctx.decode_clip(cut, n, details=True)
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
            assert got[4].tolist() == [0, 0, 1, 0, 0, 1]


def test_decode_frame_level_matches_encoder_frames():
    import basic_video_codec_b200 as bvc
    H, W, bs = 64, 96, 16
    clip = synth.moving_clip(77, H, W, 3, step=3, clamp=16)
    qps = [2, 5, 3, 4]
    for frac in (False, True):
        with bvc.Context(W, H, bs, 8, 3, 2, False, frac, 3, device=0) as ctx:
            i0 = ctx.encode_iframe(clip[0], qps)
            rec, lev, pred, q = ctx.decode_frame(True, i0.pred_bytes, i0.coef_bytes)
            assert np.array_equal(rec, i0.recon) and np.array_equal(lev, i0.levels)
            assert pred[:, 0].tolist() == i0.modes.tolist() and q.tolist() == qps
            p1 = ctx.encode_pframe(clip[1], [i0.recon], qps)
            rec, lev, pred, q = ctx.decode_frame(False, p1.pred_bytes, p1.coef_bytes, [i0.recon])
            assert np.array_equal(rec, p1.recon) and np.array_equal(lev, p1.levels) and np.array_equal(pred, p1.mv)
            p2 = ctx.encode_pframe(clip[2], [i0.recon, p1.recon])
            rec, lev, pred, q = ctx.decode_frame(False, p2.pred_bytes, p2.coef_bytes, [i0.recon, p1.recon])
            assert np.array_equal(rec, p2.recon) and np.array_equal(pred, p2.mv) and q.tolist() == [3, 3, 3, 3]
            with pytest.raises(ValueError):
                ctx.decode_frame(False, p2.pred_bytes, p2.coef_bytes, [])


def test_decode_malformed_streams_raise_value_error():
    g = gu.load("fs_i8_r4_qp3")
    e = g["meta"]["enc"]
    n, H, W = g["frames"].shape
    data = g["encoded"]
    with _ctx(W, H, e) as ctx:
        with pytest.raises(ValueError):
            ctx.decode_clip(data[:-40], n)                       # truncated record
        recs = gu.split_container(data)
        m, p, c = recs[0]
        # a coefficient payload cut in the middle: fewer EOB-terminated runs than blocks
        bad = bytes([m]) + len(p).to_bytes(2, "big") + p + (len(c) // 2).to_bytes(3, "big") + c[:len(c) // 2]
        with pytest.raises(ValueError):
            ctx.decode_clip(bad, 1)
        # prediction payload too short
        bad = bytes([m]) + (4).to_bytes(2, "big") + p[:4] + len(c).to_bytes(3, "big") + c
        with pytest.raises(ValueError):
            ctx.decode_clip(bad, 1)
        # a motion vector that leaves the plane (P frame with all-ones prediction bits = zero symbols is fine; craft a big mv)
        m1, p1, c1 = recs[1]
        evil = bytearray(p1)
        evil[0:4] = b"\x00\x00\x7f\xff"
        try:
            ctx.decode_clip(bytes([m]) + len(p).to_bytes(2, "big") + p + len(c).to_bytes(3, "big") + c +
                            bytes([m1]) + len(evil).to_bytes(2, "big") + bytes(evil) + len(c1).to_bytes(3, "big") + c1, 2)
        except ValueError:
            pass
        # the context still works afterwards
        assert np.array_equal(ctx.decode_clip(data, n), g["recon"])


@pytest.mark.parametrize("name", ["fs_i8_r4_qp3", "frac_fastme_i8_nref3", "fs_i16_r2_nref4"])
def test_decode_video_dropin(name, tmp_path):
    """decode_video(InputParameters) reads encoded.bin written by encode_video and writes mc_decoded.yuv == the
    reconstruction (what the reference's decoder logs as psnr = inf); both the clip call and the frame-object loop."""
    import basic_video_codec_b200 as bvc
    from tests.test_gpu_parity import _run_encode_video
    g = gu.load(name)
    out = _run_encode_video(tmp_path, g)
    e = g["meta"]["enc"]
    n, H, W = g["frames"].shape
    ec = bvc.EncoderConfig(e["block"], e["search_range"], e["i_period"], e["qp"], nRefFrames=e.get("nref", 1), fastME=e.get("fastme", False),
                           fracMeEnabled=e.get("frac", False), resolution=(W, H))
    params = bvc.InputParameters(str(tmp_path / "clip.y"), W, H, ec, frames_to_process=n)
    for fn in (bvc.decode_video, bvc.decode_video_framewise):
        path = os.path.join(out, "mc_decoded.yuv")
        if os.path.exists(path):
            os.unlink(path)
        fn(params)
        assert open(path, "rb").read() == g["recon"].tobytes(), fn.__name__


def test_frame_objects_decode_protocol():
    """IFrame / PFrame.entropy_decode_prediction_data -> entropy_decode_dct_coffs -> decode_mc_q_dct (decoder.py:60-69)."""
    from collections import deque
    import basic_video_codec_b200 as bvc
    from basic_video_codec_b200.encoder.IFrame import IFrame
    from basic_video_codec_b200.encoder.PFrame import PFrame
    g = gu.load("fs_i16_r2_nref4")
    e = g["meta"]["enc"]
    n, H, W = g["frames"].shape
    ec = bvc.EncoderConfig(e["block"], e["search_range"], e["i_period"], e["qp"], nRefFrames=e["nref"], resolution=(W, H))
    params = bvc.InputParameters("unused.y", W, H, ec, frames_to_process=n)
    refs = deque(maxlen=e["nref"])
    for idx, (mode, pred, coef) in enumerate(gu.split_container(g["encoded"])[:4]):
        det = g["meta"]["frames"][idx]
        fr = IFrame() if mode == 1 else PFrame(reference_frames=refs, interpolated_reference_frames=None)
        if mode == 1:
            refs.clear()
        got = fr.entropy_decode_prediction_data(pred, params)
        if mode == 1:
            assert got == det["modes"] == fr.intra_modes
        else:
            assert [list(v) for v in got.values()] == det["mv"]
            assert list(got.keys())[:2] == [(0, 0), (e["block"], 0)]
        assert fr.rc_qp_per_row == [e["qp"]] * (H // e["block"])
        fr.entropy_encoded_DCT_coffs = coef
        lev = fr.entropy_decode_dct_coffs(params)
        assert np.array_equal(lev, g["levels"][idx]) and fr.get_quat_dct_coffs_extremes() == [lev.min(), lev.max()]
        dec = fr.decode_mc_q_dct((H, W), ec)
        assert np.array_equal(dec, g["recon"][idx])
        refs.append(dec)
