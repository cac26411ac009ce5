"""Live differential tests of the C oracle against the imported Python reference (/root/reference).
They run only in the build container (marker `reference`); the committed goldens cover the GPU box."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.reference


def _rh():
    from oracle import ref_harness as rh
    return rh


class _PF:   # the two attributes block_predictor reads off the frame object
    def __init__(self, refs, irefs):
        self.reference_frames, self.interpolated_reference_frames = refs, irefs


@pytest.mark.parametrize("frac", [False, True])
@pytest.mark.parametrize("content", ["moving", "poster"])
def test_full_search_block_by_block(frac, content):
    """find_lowest_mae_block (block_predictor.py:61-91) vs bvo_full_search_block on corner / edge / interior
    blocks, 1-3 references, tie-heavy content: MV, reference index and SAD."""
    rh = _rh()
    from oracle import bindings as ob
    ns = rh.load_reference()
    H, W, bs, r = 48, 64, 8, 3
    clip = (synth.moving_clip(31, H, W, 4, step=2, clamp=8) if content == "moving" else synth.posterised_clip(32, H, W, 4))
    for nref in (1, 3):
        refs = [clip[i] for i in range(nref)]
        cur = clip[3]
        ec = rh.make_config(ns, block=bs, search_range=r, qp=3, i_period=8, nref=nref, frac=frac, width=W, height=H)
        irefs = [ns.block_predictor.build_pre_interpolated_buffer(x) if frac else None for x in refs]
        cfg = ob.make_config(W, H, bs, r, 3, nref=nref, frac=frac)
        mv_o, sad_o, _ = ob.me_frame(cfg, cur, [ob.halfpel_plane(x) for x in refs] if frac else refs)
        bw = W // bs
        for (bx, by) in [(0, 0), (bw - 1, 0), (0, H // bs - 1), (bw - 1, H // bs - 1), (3, 2), (1, 4), (bw - 2, 1)]:
            blk = cur[by * bs:(by + 1) * bs, bx * bs:(bx + 1) * bs].astype(np.int16)
            mv, mae, _, _ = ns.block_predictor.find_lowest_mae_block(blk, (bx * bs, by * bs), _PF(refs, irefs), ec)
            b = by * bw + bx
            assert list(mv) == mv_o[b].tolist()
            assert mae * bs * bs == sad_o[b]


@pytest.mark.parametrize("frac", [False, True])
def test_fast_me_block_by_block(frac):
    """find_fast_me_block (block_predictor.py:11-58) incl. the late-binding closure behaviour, random MVPs."""
    rh = _rh()
    from oracle import bindings as ob
    import ctypes as C
    ns = rh.load_reference()
    H, W, bs = 48, 64, 8
    clip = synth.moving_clip(33, H, W, 5, step=3, clamp=8, blur=15)
    rng = np.random.default_rng(3)
    L = ob.lib()
    for nref in (1, 2, 4):
        refs = [clip[i] for i in range(nref)]
        cur = np.ascontiguousarray(clip[4])
        ec = rh.make_config(ns, block=bs, search_range=4, qp=3, i_period=8, nref=nref, fastme=True, frac=frac, width=W, height=H)
        irefs = [ns.block_predictor.build_pre_interpolated_buffer(x) if frac else None for x in refs]
        planes = [ob.halfpel_plane(x) for x in refs] if frac else [np.ascontiguousarray(x) for x in refs]
        arr = (C.c_void_p * nref)(*[p.ctypes.data for p in planes])
        for _ in range(12):
            bx, by = int(rng.integers(0, W // bs)), int(rng.integers(0, H // bs))
            mvp = (int(rng.integers(-5, 6)), int(rng.integers(-5, 6)))
            blk = cur[by * bs:(by + 1) * bs, bx * bs:(bx + 1) * bs].astype(np.int16)
            mv, mae, _, cnt = ns.block_predictor.find_fast_me_block(blk, (bx * bs, by * bs), mvp, _PF(refs, irefs), ec, 0)
            m = (C.c_int32 * 3)()
            cmp_ = C.c_int64(0)
            s = L.bvo_fast_me_block(cur.ctypes.data_as(C.c_void_p), W, H, bx * bs, by * bs, bs, arr, nref, int(frac), mvp[0], mvp[1], m, C.byref(cmp_))
            assert tuple(mv) == (m[0], m[1], m[2])
            assert mae * bs * bs == s and cnt == cmp_.value


def test_dct_against_scipy_fp64_and_tie_statistics():
    """North-star tolerance: pre-quantisation coefficients within 1e-9 of the reference's SciPy DCT run on
    float64.  Levels may differ only where coef/Q sits on an exact tie -- counted, not hidden."""
    rh = _rh()
    from oracle import bindings as ob
    ns = rh.load_reference()
    rh.set_dct_mode("fp64_scipy")
    rng = np.random.default_rng(9)
    try:
        for bs, qp in ((8, 3), (16, 3), (4, 0)):
            mism = ties = total = 0
            for _ in range(400):
                res = rng.integers(-255, 256, size=(bs, bs)).astype(np.int16)
                ref_coef = ns.dct.apply_dct_2d(res)
                coef = ob.fdct(res)
                assert np.max(np.abs(coef - ref_coef)) < 1e-9
                Q = ns.dct.generate_quantization_matrix(bs, qp)
                ref_lev = ns.dct.quantize_block(ref_coef, Q)
                lev, _, _, _ = ob.transform_block(res, np.zeros((bs, bs), np.int16), qp)
                d = lev != ref_lev
                frac = np.abs(coef / Q - np.floor(coef / Q) - 0.5)
                assert np.all(frac[d] < 1e-9), "a level differs away from an exact quantiser tie"
                mism += int(d.sum())
                ties += int((frac < 1e-9).sum())
                total += bs * bs
            print(f"bs={bs} qp={qp}: {mism} level mismatches, all on {ties} exact ties, of {total} coefficients")
    finally:
        rh.set_dct_mode("fp64_defined")


def test_reference_encode_video_equals_oracle_clip():
    """The reference's own encode_video (defined DCT patched in) against the oracle's clip encoder, live."""
    rh = _rh()
    from oracle import bindings as ob
    frames = synth.moving_clip(41, 32, 48, 5, step=2, clamp=8)
    out = rh.ref_encode_video(frames, block=8, search_range=2, qp=2, i_period=3, nref=2)
    data, recon = ob.encode_clip(ob.make_config(48, 32, 8, 2, 2, nref=2, i_period=3), frames)
    assert data == out["encoded"] and np.array_equal(recon, out["recon"])


def test_whole_clip_divergence_from_the_shipped_float32_dct_is_recorded_and_only_ties():
    """The drop-in contract is "the reference with the defined fp64 DCT".  tests/golden/cif_c1_dct_divergence.json
    (oracle/gen_dct_divergence.py) records how far the reference as shipped (float32 SciPy) and with SciPy on float64 are
    from it on the CIF config-1 stand-in.  Here the first two frames are re-encoded with the imported reference and must
    reproduce the recorded per-frame numbers; and the same-residual probe must find differing levels only at exact
    quantiser ties (defined coefficient / Q exactly k + 1/2, SciPy within 1e-9 of it)."""
    import json
    import os
    _rh()
    from oracle import gen_dct_divergence as gd
    from tests import golden_util as gu
    rec = json.load(open(os.path.join(gu.GOLD, "cif_c1_dct_divergence.json")))
    live = gd.run(2)
    for mode_live, mode_rec in zip(live["modes"], rec["modes"]):
        assert mode_live["mode"] == mode_rec["mode"]
        for fl, fr in zip(mode_live["frames"], mode_rec["frames"][:2]):
            assert fl == fr
    probe = live["same_residual_probe_fp64_scipy_vs_defined"]
    assert probe["blocks"] > 0 and probe["non_tie_diffs"] == 0 and probe["max_tie_distance"] == 0.0
    assert probe["max_coef_diff_all"] < 1e-9
