"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed goldens.
Bit-exact for every integer / byte / index output; coefficients within 1e-9 (they are in fact bit-equal)."""
import hashlib

import numpy as np
import pytest

from tests import golden_util as gu
from tests import synth

pytestmark = pytest.mark.gpu


def _ob():
    from oracle import bindings as ob
    return ob


def _ctx(W, H, bs, r, qp, nref=1, fastme=False, frac=False, ip=1, lanes=1):
    import basic_video_codec_b200 as bvc
    return bvc.Context(W, H, bs, r, qp, nref, fastme, frac, ip, device=0, max_lanes=lanes)


# ---- K1/K3/K4 motion estimation -----------------------------------------------------------------
ME_CASES = [
    # (H, W, bs, r, nref, frac, fastme, content)
    (64, 96, 16, 8, 1, False, False, "moving"),     # tiled, FIRST+LAST bodies only
    (96, 128, 16, 16, 2, False, False, "moving"),   # tiled with one MID body, 2 refs
    (128, 192, 16, 32, 1, False, False, "moving"),  # the headline configuration's kernel
    (64, 96, 8, 4, 1, False, False, "moving"),      # config-1 kernel
    (64, 96, 8, 8, 3, False, False, "poster"),      # ties, 3 refs
    (32, 48, 4, 2, 2, False, False, "poster"),
    (64, 96, 16, 2, 4, False, False, "moving"),     # generic kernel (2R < bs)
    (48, 80, 8, 3, 2, False, False, "poster"),      # generic kernel (2R % bs != 0)
    (64, 96, 8, 4, 2, True, False, "moving"),       # half-pel, tiled on phase planes
    (64, 96, 16, 8, 1, True, False, "poster"),      # half-pel ties
    (48, 64, 8, 2, 2, True, False, "moving"),       # half-pel generic
    (64, 96, 16, 4, 4, False, True, "moving"),      # FastME 4 refs
    (64, 96, 8, 4, 3, True, True, "moving"),        # FastME half-pel
    (64, 96, 16, 4, 2, False, True, "poster"),      # FastME ties
]


@pytest.mark.parametrize("H,W,bs,r,nref,frac,fastme,content", ME_CASES)
def test_me_matches_oracle(H, W, bs, r, nref, frac, fastme, content):
    ob = _ob()
    n = nref + 1
    clip = (synth.moving_clip(100 + bs + r, H, W, n, step=min(r, 6), clamp=24) if content == "moving"
            else synth.posterised_clip(200 + bs + r, H, W, n))
    cur, refs = clip[-1], [clip[i] for i in range(nref)]
    cfg = ob.make_config(W, H, bs, r, 3, nref=nref, fastme=fastme, frac=frac)
    planes = [ob.halfpel_plane(x) for x in refs] if frac else refs
    mv_o, sad_o, cmp_o = ob.me_frame(cfg, cur, planes)
    with _ctx(W, H, bs, r, 3, nref, fastme, frac) as ctx:
        mv_g, sad_g, cmp_g = ctx.me_search(cur, refs)
    assert np.array_equal(sad_g, sad_o)
    assert np.array_equal(mv_g, mv_o)
    assert cmp_g == cmp_o


def test_me_flat_frame_all_ties():
    """Every candidate has SAD 0: the winner must be (0,0,0) for every block (min L1, first ref)."""
    ob = _ob()
    H, W, bs, r = 64, 96, 16, 8
    cur = np.full((H, W), 77, np.uint8)
    refs = [cur.copy(), cur.copy()]
    with _ctx(W, H, bs, r, 3, 2) as ctx:
        mv, sad, _ = ctx.me_search(cur, refs)
    assert not mv.any() and not sad.any()


# ---- K2 half-pel ------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W", [(32, 48), (64, 96), (288, 352)])
def test_halfpel_matches_oracle(H, W):
    ob = _ob()
    ref = synth.texture(5, H, W, blur=3)
    with _ctx(W, H, 8, 2, 3, 1, False, True) as ctx:
        got = ctx.interp_halfpel(ref)
    assert np.array_equal(got, ob.halfpel_plane(ref))


# ---- K5 transform ------------------------------------------------------------------------------------
@pytest.mark.parametrize("bs", [4, 8, 16])
@pytest.mark.parametrize("qp", [0, 3, 6])
def test_transform_blocks_match_oracle(bs, qp):
    from basic_video_codec_b200._lib import dct_quant_recon
    ob = _ob()
    rng = np.random.default_rng(bs * 10 + qp)
    n = 203
    res = rng.integers(-255, 256, size=(n, bs, bs)).astype(np.int16)
    res[0] = 0
    res[1] = 255
    res[2] = -255
    res[3] = (rng.integers(-4, 5, size=(bs, bs)) * 64).astype(np.int16)   # posterised residuals: exact ties
    res[4, :, :] = 0
    res[4, 0, 0] = bs * 4 * 3                                             # DC = 12 -> tie at qp 3
    pred = rng.integers(0, 256, size=(n, bs, bs)).astype(np.int16)
    level, recon, idct, coef = dct_quant_recon(res, pred, qp)
    for i in range(n):
        l_o, r_o, i_o, c_o = ob.transform_block(res[i], pred[i], qp)
        assert np.max(np.abs(coef[i] - c_o)) < 1e-9        # north-star tolerance on pre-quantisation coefficients
        assert np.array_equal(coef[i], c_o)                # and in fact bit-identical
        assert np.array_equal(level[i], l_o)
        assert np.array_equal(recon[i], r_o)
        assert np.array_equal(idct[i], i_o)


# ---- frame level -------------------------------------------------------------------------------------
FRAME_CASES = [
    dict(H=64, W=96, bs=8, r=4, qp=3, nref=1, frac=False, fastme=False),
    dict(H=64, W=96, bs=16, r=8, qp=2, nref=3, frac=False, fastme=False),
    dict(H=48, W=64, bs=8, r=2, qp=1, nref=2, frac=True, fastme=False),
    dict(H=64, W=96, bs=16, r=4, qp=4, nref=2, frac=False, fastme=True),
    dict(H=32, W=48, bs=4, r=2, qp=0, nref=2, frac=False, fastme=False),
    dict(H=64, W=96, bs=16, r=8, qp=9, nref=1, frac=False, fastme=False),
]


@pytest.mark.parametrize("case", FRAME_CASES)
def test_frames_match_oracle(case):
    ob = _ob()
    H, W, bs, r, qp, nref = case["H"], case["W"], case["bs"], case["r"], case["qp"], case["nref"]
    frac, fastme = case["frac"], case["fastme"]
    clip = synth.moving_clip(300 + bs + qp, H, W, nref + 2, step=3, clamp=16)
    cfg = ob.make_config(W, H, bs, r, qp, nref=nref, fastme=fastme, frac=frac)
    rows = H // bs
    qp_rows = np.array([max(0, qp + (i % 3) - 1) for i in range(rows)], np.int32)   # per-row QPs (RC hook)
    with _ctx(W, H, bs, r, qp, nref, fastme, frac) as ctx:
        refs_o, refs_g = [], []
        for idx in range(clip.shape[0]):
            qr = qp_rows if idx % 2 else None
            if idx == 0:
                o = ob.encode_iframe(cfg, clip[idx], qr)
                g = ctx.encode_iframe(clip[idx], qr)
                assert np.array_equal(g.modes, o.modes)
                assert np.array_equal(g.resid_mc.view(np.uint8), o.resid_mc.view(np.uint8))
            else:
                hp = [ob.halfpel_plane(x) for x in refs_o] if frac else None
                o = ob.encode_pframe(cfg, clip[idx], refs_o, hp, qr)
                g = ctx.encode_pframe(clip[idx], refs_g, qr)
                assert np.array_equal(g.mv, o.mv)
                assert np.array_equal(g.resid_mc, o.resid_mc)
                assert np.array_equal(g.resid_nomc, o.resid_nomc)
            assert np.array_equal(g.sad, o.sad)
            assert np.array_equal(g.levels, o.levels)
            assert np.array_equal(g.recon, o.recon)
            assert (g.pred_nbits, g.coef_nbits) == (o.pred_nbits, o.coef_nbits)
            assert g.pred_bytes == o.pred_bytes
            assert g.coef_bytes == o.coef_bytes
            assert g.bits_per_row.tolist() == o.bits_per_row.tolist()
            assert g.avg_mae == o.avg_mae
            assert g.mae_comparisons == o.mae_comparisons
            refs_o.append(o.recon)
            refs_g.append(g.recon)
            if len(refs_o) > nref:
                refs_o.pop(0)
                refs_g.pop(0)


# ---- clip level: goldens from the Python reference -----------------------------------------------------
@pytest.mark.parametrize("name", [n for n in gu.names() if not n.startswith("rc")])
def test_clip_matches_reference_golden(name):
    g = gu.load(name)
    frames, e = g["frames"], g["meta"]["enc"]
    n, H, W = frames.shape
    for lanes in (1, 3):
        with _ctx(W, H, e["block"], e["search_range"], e["qp"], e.get("nref", 1), e.get("fastme", False),
                  e.get("frac", False), e["i_period"], lanes=lanes) as ctx:
            data, recon = ctx.encode_clip(frames, want_recon=True)
        assert hashlib.sha256(data).hexdigest() == g["meta"]["encoded_sha256"], f"lanes={lanes}"
        if "recon" in g:
            assert np.array_equal(recon, g["recon"])
        else:
            assert hashlib.sha256(recon.tobytes()).hexdigest() == g["meta"]["recon_sha256"]


def test_clip_many_gops_resident_matches_oracle():
    """GOP lanes + waves + a short last GOP, inputs resident in HBM, against the oracle's clip encoder."""
    ob = _ob()
    H, W, bs, r, qp, ip, n = 64, 96, 16, 8, 3, 4, 23
    frames = synth.moving_clip(7, H, W, n, step=4, clamp=24)
    cfg = ob.make_config(W, H, bs, r, qp, nref=2, i_period=ip)
    want, _ = ob.encode_clip(cfg, frames, want_recon=False)
    with _ctx(W, H, bs, r, qp, 2, False, False, ip, lanes=4) as ctx:
        ctx.clip_upload(frames)
        out, ln = ctx.encode_clip_resident(n)
        assert out[:ln].tobytes() == want
        # idempotence: a second pass over the resident clip gives the same stream
        out2, ln2 = ctx.encode_clip_resident(n)
        assert ln2 == ln and np.array_equal(out2[:ln2], out[:ln])
        assert ctx.launch_count() > 0


def test_gop_streams_concatenate():
    """Per-GOP streams concatenate to the whole-clip stream (the property GOP sharding rests on)."""
    H, W, bs, r, qp, ip, n = 64, 96, 8, 4, 3, 3, 9
    frames = synth.moving_clip(9, H, W, n, step=3, clamp=16)
    with _ctx(W, H, bs, r, qp, 1, False, False, ip, lanes=3) as ctx:
        whole, _ = ctx.encode_clip(frames)
        parts = b"".join(ctx.encode_clip(frames[g * ip:(g + 1) * ip])[0] for g in range(n // ip))
    assert whole == parts


def test_full_size_1080p_properties():
    """BASELINE config 4 geometry (1920x1088, i=16, r=32) on a 2-GOP sample: the stream parses, the ME
    of a frame against itself is all-zero, and the oracle agrees on a sampled block row."""
    ob = _ob()
    H, W, bs, r, qp = 1088, 1920, 16, 32, 4
    frames = synth.moving_clip(1080, H, W, 3, step=6, clamp=96)
    with _ctx(W, H, bs, r, qp, 1, False, False, 30, lanes=1) as ctx:
        mv, sad, _ = ctx.me_search(frames[1], [frames[1]])
        assert not mv.any() and not sad.any()
        mv, sad, _ = ctx.me_search(frames[1], [frames[0]])
        # oracle on the first and last block rows and one interior row (borders + interior), block by block
        L = ob.lib()
        import ctypes as C
        ref = np.ascontiguousarray(frames[0])
        arr = (C.c_void_p * 1)(ref.ctypes.data)
        cur = np.ascontiguousarray(frames[1])
        bw = W // bs
        for by in (0, 33, H // bs - 1):
            for bx in list(range(0, bw, 17)) + [bw - 1]:
                m = (C.c_int32 * 3)()
                s = L.bvo_full_search_block(cur.ctypes.data_as(C.c_void_p), W, H, bx * bs, by * bs, bs, arr, 1, r, 0, m, None)
                b = by * bw + bx
                assert (mv[b, 0], mv[b, 1], mv[b, 2], sad[b]) == (m[0], m[1], m[2], s), (bx, by)
        data, recon = ctx.encode_clip(frames, want_recon=True)
        parts = gu.split_container(data)
        assert [p[0] for p in parts] == [1, 0, 0]
        assert recon.shape == frames.shape
        assert ctx.me_work_per_frame(1) == 8527896576


# ---- Python drop-in layer --------------------------------------------------------------------------------
def _run_encode_video(tmp_path, g):
    from basic_video_codec_b200 import EncoderConfig, InputParameters
    from basic_video_codec_b200.encoder.encoder import encode_video, output_dir
    frames, e = g["frames"], g["meta"]["enc"]
    n, H, W = frames.shape
    yfile = tmp_path / "clip.y"
    yfile.write_bytes(frames.tobytes())
    ec = EncoderConfig(e["block"], e["search_range"], e["i_period"], e["qp"], nRefFrames=e.get("nref", 1),
                       fastME=e.get("fastme", False), fracMeEnabled=e.get("frac", False), resolution=(W, H))
    params = InputParameters(str(yfile), W, H, ec, frames_to_process=n)
    encode_video(params)
    return output_dir(params)


@pytest.mark.parametrize("name", ["fs_i8_r4_qp3", "fastme_i16_nref4", "frac_fs_i8_r2_nref2", "fs_i16_r2_nref4"])
def test_encode_video_dropin_writes_reference_files(name, tmp_path):
    """encode_video(InputParameters) through PFrame/IFrame objects: every side file the reference writes
    (encoder.py:104-152, file_io.py) is byte-identical to the reference's own output."""
    import os
    g = gu.load(name)
    out = _run_encode_video(tmp_path, g)
    e = g["meta"]["enc"]
    sr = -1 if e.get("fastme") else e["search_range"]
    ident = f'{e["block"]}_{sr}{".0" if e.get("frac") else ""}_{e["qp"]}_{e["i_period"]}_{e.get("nref", 1)}_0_0'
    assert out.endswith(os.path.join("clip", ident))          # output naming scheme, file_io.py:20-26
    rd = lambda f: open(os.path.join(out, f), "rb").read()
    assert rd("encoded.bin") == g["encoded"]
    assert rd("mc_reconstructed.yuv") == g["recon"].tobytes()
    assert rd("mc_quant_dct_coff.bin") == g["levels"].tobytes()
    assert rd("residuals_w_mc.yuv") == g["resid_mc"].tobytes()
    assert rd("residuals_wo_mc.yuv") == g["resid_nomc"].tobytes()
    assert open(os.path.join(out, "mv.txt")).read() == g["mv_txt"]
    rows = open(os.path.join(out, "metrics.csv")).read().strip().splitlines()
    assert rows[0] == "idx,I-Frame,avg_MAE,mae_comps,PSNR,frame_bytes,file_bits,enc_time,elapsed_time"
    assert len(rows) == 1 + g["frames"].shape[0]
    for r, det in zip(rows[1:], g["meta"]["frames"]):
        f = r.split(",")
        assert int(f[1]) == det["intra"] and float(f[2]) == det["avg_mae"] and int(f[3]) == det["mae_comparisons"]


def test_frame_objects_keep_reference_attribute_protocol():
    from collections import deque
    from basic_video_codec_b200 import EncoderConfig
    from basic_video_codec_b200.encoder.IFrame import IFrame
    from basic_video_codec_b200.encoder.PFrame import PFrame
    from basic_video_codec_b200.encoder.PredictionMode import PredictionMode
    g = gu.load("fs_i16_r2_nref4")
    frames, e = g["frames"], g["meta"]["enc"]
    ec = EncoderConfig(e["block"], e["search_range"], e["i_period"], e["qp"], nRefFrames=e["nref"], resolution=frames.shape[:0:-1])
    refs = deque(maxlen=e["nref"])
    fi = IFrame(frames[0])
    fi.encode_mc_q_dct(ec)
    assert fi.prediction_mode == PredictionMode.INTRA_FRAME and fi.is_iframe()
    assert fi.intra_modes == g["meta"]["frames"][0]["modes"]
    assert len(fi.entropy_encoded_DCT_coffs) == g["meta"]["frames"][0]["coef_nbits"]
    refs.append(fi.reconstructed_frame)
    fp = PFrame(frames[1], refs, deque())
    fp.encode_mc_q_dct(ec)
    det = g["meta"]["frames"][1]
    assert fp.is_pframe() and list(fp.mv_field.values()) == det["mv"]
    assert list(fp.mv_field.keys())[:2] == [(0, 0), (16, 0)]
    assert fp.bits_per_row == det["bits_per_row"] and fp.total_mae_comparisons == det["mae_comparisons"]
    assert fp.get_mv_extremes()[0] == np.array(det["mv"]).min(axis=0).tolist()
    assert fp.get_quat_dct_coffs_extremes() == [g["levels"][1].min(), g["levels"][1].max()]


# ---- edge cases ------------------------------------------------------------------------------------------
EDGE = [
    # (H, W, bs, r, qp, nref, ip, n, frac, fastme)
    (16, 16, 16, 8, 3, 1, 2, 3, False, False),     # a single block: every non-zero candidate leaves the frame
    (16, 48, 16, 32, 3, 2, 3, 4, False, False),    # search range larger than the frame
    (24, 40, 8, 4, 0, 3, 1, 3, False, False),      # I_Period 1 (all intra), qp 0, width not a multiple of 16
    (32, 32, 4, 2, 9, 2, 4, 5, False, False),      # qp at the validation limit log2(4)+7
    (32, 48, 8, 4, 10, 1, 5, 5, True, False),      # half-pel at max qp
    (32, 48, 16, 4, 5, 4, 8, 7, True, True),       # FastME half-pel, window longer than the clip start
    (40, 56, 8, 8, 2, 8, 9, 9, False, False),      # 8 reference frames
]


@pytest.mark.parametrize("H,W,bs,r,qp,nref,ip,n,frac,fastme", EDGE)
def test_edge_cases_match_oracle(H, W, bs, r, qp, nref, ip, n, frac, fastme):
    ob = _ob()
    frames = synth.moving_clip(500 + H + W + bs, H, W, n, step=3, clamp=8, blur=3)
    cfg = ob.make_config(W, H, bs, r, qp, nref=nref, fastme=fastme, frac=frac, i_period=ip)
    want, want_recon = ob.encode_clip(cfg, frames)
    with _ctx(W, H, bs, r, qp, nref, fastme, frac, ip, lanes=2) as ctx:
        got, recon = ctx.encode_clip(frames, want_recon=True)
    assert np.array_equal(recon, want_recon)
    assert got == want


def test_extreme_content():
    """Saturated / flat / checkerboard planes: clipping in reconstruction, all-zero blocks, maximal levels."""
    ob = _ob()
    H, W, bs = 32, 48, 8
    rng = np.random.default_rng(5)
    planes = [np.zeros((H, W), np.uint8), np.full((H, W), 255, np.uint8),
              ((np.indices((H, W)).sum(0) % 2) * 255).astype(np.uint8), rng.integers(0, 256, (H, W)).astype(np.uint8),
              np.full((H, W), 255, np.uint8), np.zeros((H, W), np.uint8)]
    frames = np.stack(planes)
    for qp in (0, 5):
        cfg = ob.make_config(W, H, bs, 4, qp, nref=2, i_period=6)
        want, want_recon = ob.encode_clip(cfg, frames)
        with _ctx(W, H, bs, 4, qp, 2, False, False, 6) as ctx:
            got, recon = ctx.encode_clip(frames, want_recon=True)
        assert got == want and np.array_equal(recon, want_recon)


def test_error_mapping():
    import basic_video_codec_b200 as bvc
    with pytest.raises(ValueError):
        bvc.Context(64, 64, 8, 4, 11)          # qp > log2(8)+7  (params.py:29-30)
    with pytest.raises(ValueError):
        bvc.Context(8, 8, 16, 4, 3)            # frame smaller than a block (block_predictor.py:70-71)
    with pytest.raises(NotImplementedError):
        bvc.Context(60, 64, 8, 4, 3)           # not a multiple of the block size: pad first
    with _ctx(32, 32, 8, 4, 3) as ctx:
        with pytest.raises(ValueError):
            ctx.encode_pframe(np.zeros((32, 32), np.uint8), [])       # empty reference window
        with pytest.raises(MemoryError):
            ctx.encode_clip(np.zeros((2, 32, 32), np.uint8), out_capacity=4)


def test_lane_groups_do_not_change_the_stream():
    """bvc_set_lane_groups only changes which CUDA streams the lanes of a step run on: same bytes for 1..4 groups,
    including a short last GOP, more groups than lanes, and half-pel planes (which add a kernel to the post chain)."""
    ob = _ob()
    H, W, bs, r, qp, ip, n = 64, 96, 16, 8, 3, 4, 27
    frames = synth.moving_clip(21, H, W, n, step=4, clamp=24)
    for frac, nref in ((False, 2), (True, 1)):
        cfg = ob.make_config(W, H, bs, r, qp, nref=nref, frac=frac, i_period=ip)
        want, want_recon = ob.encode_clip(cfg, frames)
        for lanes, groups in ((7, 1), (7, 2), (7, 3), (7, 4), (2, 4), (1, 2)):
            with _ctx(W, H, bs, r, qp, nref, False, frac, ip, lanes=lanes) as ctx:
                ctx.set_lane_groups(groups)
                data, recon = ctx.encode_clip(frames, want_recon=True)
            assert data == want, f"frac={frac} lanes={lanes} groups={groups}"
            assert np.array_equal(recon, want_recon)
    with _ctx(W, H, bs, r, qp, 1, False, False, ip, lanes=2) as ctx:
        with pytest.raises(ValueError):
            ctx.set_lane_groups(0)
        with pytest.raises(ValueError):
            ctx.set_lane_groups(5)


@pytest.mark.parametrize("frac,nref", [(False, 3), (True, 2)])
def test_fastme_clip_with_lane_groups_and_modes(frac, nref):
    """FastME through the clip call: 15 blocks per frame (an odd count: the per-group slices of the transfer-table scratch
    must stay 16-byte aligned), 1..3 lane groups running on their own streams, every evaluation mode: same bytes as
    the oracle."""
    ob = _ob()
    H, W, bs, qp, ip, n = 48, 80, 16, 3, 4, 23
    frames = synth.moving_clip(33, H, W, n, step=5, clamp=24, blur=5)
    cfg = ob.make_config(W, H, bs, 4, qp, nref=nref, fastme=True, frac=frac, i_period=ip)
    want, want_recon = ob.encode_clip(cfg, frames)
    for lanes, groups, mode in ((6, 1, 0), (6, 2, 4), (5, 3, 4), (6, 2, 2), (3, 2, 1), (6, 2, 3), (5, 3, 3)):
        with _ctx(W, H, bs, 4, qp, nref, True, frac, ip, lanes=lanes) as ctx:
            ctx.set_lane_groups(groups)
            ctx.set_fastme_direct(mode)
            data, recon = ctx.encode_clip(frames, want_recon=True)
        assert data == want, f"lanes={lanes} groups={groups} mode={mode}"
        assert np.array_equal(recon, want_recon)


@pytest.mark.parametrize("H,W,bs", [(16, 16, 16), (48, 16, 16), (16, 64, 16), (8, 40, 8), (36, 36, 4)])
def test_fastme_modes_on_degenerate_geometry(H, W, bs):
    """A single block, a single block column / row, and block sizes below 16: every candidate but the origin leaves the
    plane in at least one direction, the half-pel phases shrink the valid range by one more pixel; every FastME
    evaluation mode against the oracle."""
    ob = _ob()
    frames = synth.moving_clip(700 + H + W, H, W, 5, step=2, clamp=6, blur=3)
    for frac in (False, True):
        cfg = ob.make_config(W, H, bs, 4, 2, nref=2, fastme=True, frac=frac, i_period=5)
        want, want_recon = ob.encode_clip(cfg, frames)
        for mode in (0, 1, 2, 3, 4):
            with _ctx(W, H, bs, 4, 2, 2, True, frac, 5, lanes=1) as ctx:
                ctx.set_fastme_direct(mode)
                data, recon = ctx.encode_clip(frames, want_recon=True)
            assert data == want, f"frac={frac} mode={mode}"
            assert np.array_equal(recon, want_recon)


def test_i420_input_stage_pads_and_skips_chroma(tmp_path):
    """bvc_clip_upload_i420: luma planes straight from an I420 file image, padded with 128 on the device, encode to the
    same stream as the oracle fed with pad_frame()'d Y planes; aligned sizes take the single strided copy."""
    from basic_video_codec_b200 import EncoderConfig
    from basic_video_codec_b200 import input_stage as ist
    ob = _ob()
    rng = np.random.default_rng(11)
    for (w, h, bs) in ((90, 58, 8), (96, 64, 16)):
        n, ip, qp, r = 7, 3, 3, 4
        W, H = w + (-w) % bs, h + (-h) % bs
        ys = synth.moving_clip(13, H, W, n, step=3, clamp=16)[:, :h, :w]
        path = tmp_path / f"c_{w}.yuv"
        with open(path, "wb") as fh:
            for y in ys:
                fh.write(np.ascontiguousarray(y).tobytes())
                fh.write(rng.integers(0, 256, 2 * (w // 2) * (h // 2), dtype=np.uint8).tobytes())
        padded = np.stack([ist.pad_frame(y, bs) for y in ys])
        want, _ = ob.encode_clip(ob.make_config(W, H, bs, r, qp, nref=2, i_period=ip), padded)
        ec = EncoderConfig(bs, r, ip, qp, nRefFrames=2, resolution=(W, H))
        assert ist.encode_yuv_file(str(path), w, h, ec) == want
        assert ist.encode_yuv_file(str(path), w, h, ec, frames_to_process=4) == ob.encode_clip(ob.make_config(W, H, bs, r, qp, nref=2, i_period=ip), padded[:4])[0]
    with _ctx(96, 64, 16, 4, 3, 1, False, False, 3, lanes=1) as ctx:
        with pytest.raises(ValueError):
            ctx.clip_upload_i420(np.zeros(100, np.uint8), 96, 64, 2)          # buffer too small
        with pytest.raises(ValueError):
            ctx.clip_upload_i420(np.zeros(10 ** 5, np.uint8), 64, 64, 2)      # not the context's size


@pytest.mark.parametrize("frac,bs,nref", [(False, 16, 2), (True, 8, 3), (False, 8, 1), (False, 4, 6), (False, 16, 6), (True, 16, 5), (True, 16, 1)])
def test_fastme_sad_map_and_direct_paths_agree_with_oracle(frac, bs, nref):
    """FastME from the SAD map (default) and with direct evaluation (bvc_set_fastme_direct) against the oracle, on
    content whose motion exceeds the +-16 MV-unit map: a smooth gradient shifted 21 pixels makes the predictor drift
    past 16 from block to block, so the walk leaves the map and the warp-evaluated fallback is exercised."""
    ob = _ob()
    H, W = 96, 160
    yy, xx = np.mgrid[0:H, 0:W + 64]
    base = ((np.sin(xx / 23.0) * 0.5 + 0.5) * 180 + (yy % 32) * 2).astype(np.uint8)      # smooth in x: SAD falls monotonically
    rng = np.random.default_rng(9)
    refs = [np.ascontiguousarray(base[:, 21 + 2 * i: 21 + 2 * i + W]) for i in range(nref)]
    cur = np.clip(base[:, :W].astype(np.int16) + rng.integers(-1, 2, (H, W)), 0, 255).astype(np.uint8)
    cfg = ob.make_config(W, H, bs, 4, 3, nref=nref, fastme=True, frac=frac)
    planes = [ob.halfpel_plane(x) for x in refs] if frac else refs
    mv_o, sad_o, cmp_o = ob.me_frame(cfg, cur, planes)
    assert np.abs(mv_o[:, 0]).max() > 16, "the case must leave the SAD map (+-16 MV units)"
    for direct in (0, 1, 2, 3, 4):  # auto, direct evaluation, serial walk on the SAD map, window walk, transfer tables
        with _ctx(W, H, bs, 4, 3, nref, True, frac) as ctx:
            ctx.set_fastme_direct(direct)
            mv_g, sad_g, cmp_g = ctx.me_search(cur, refs)
        assert np.array_equal(mv_g, mv_o) and np.array_equal(sad_g, sad_o) and cmp_g == cmp_o, f"direct={direct}"


def test_two_devices_in_one_process():
    """One process, one context per GPU (host threads / GOP sharding without torchrun): kernel attributes are configured
    per device, so the second device must work as well as the first."""
    import ctypes
    import basic_video_codec_b200 as bvc
    try:
        n = ctypes.c_int(0)
        ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n))
        ndev = n.value
    except OSError:
        import torch
        ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs two GPUs")
    ob = _ob()
    H, W, bs, r, qp, ip, nfr = 128, 192, 16, 32, 3, 3, 6     # r = 32: the tiled kernel with > 48 KB of shared memory
    frames = synth.moving_clip(41, H, W, nfr, step=4, clamp=24)
    want, _ = ob.encode_clip(ob.make_config(W, H, bs, r, qp, nref=2, i_period=ip), frames, want_recon=False)
    for dev in (0, 1):
        with bvc.Context(W, H, bs, r, qp, 2, False, False, ip, device=dev, max_lanes=2) as ctx:
            data, recon = ctx.encode_clip(frames, want_recon=True)
            assert data == want, f"device {dev}"
            assert np.array_equal(ctx.decode_clip(data, nfr), recon)
        with bvc.Context(W, H, bs, 4, qp, 2, True, False, ip, device=dev, max_lanes=2) as ctx:
            assert ctx.encode_clip(frames)[0] == ob.encode_clip(ob.make_config(W, H, bs, 4, qp, nref=2, fastme=True, i_period=ip), frames, want_recon=False)[0]


def test_dense_blocks_bs16_qp0():
    """16x16 blocks with every level non-zero and near the largest magnitude (white noise / checkerboard at qp 0): the
    entropy coder's dense path (more than 32 symbol positions per block), the per-block bit buffer at its worst case and
    the container's length fields."""
    ob = _ob()
    H, W, bs = 64, 96, 16
    rng = np.random.default_rng(17)
    chk = ((np.indices((H, W)).sum(0) % 2) * 255).astype(np.uint8)
    frames = np.stack([rng.integers(0, 256, (H, W)).astype(np.uint8), chk, rng.integers(0, 256, (H, W)).astype(np.uint8),
                       255 - chk, rng.integers(0, 2, (H, W)).astype(np.uint8) * 255])
    for qp, ip in ((0, 5), (0, 1), (2, 3)):
        cfg = ob.make_config(W, H, bs, 4, qp, nref=1, i_period=ip)
        want, want_recon = ob.encode_clip(cfg, frames)
        with _ctx(W, H, bs, 4, qp, 1, False, False, ip, lanes=2) as ctx:
            got, recon = ctx.encode_clip(frames, want_recon=True)
            assert got == want and np.array_equal(recon, want_recon), (qp, ip)
            assert np.array_equal(ctx.decode_clip(got, len(frames)), recon)
    bits_per_block = len(want) * 8 / (len(frames) * (H // bs) * (W // bs))
    assert bits_per_block > 600      # the case really is dense
