"""The C oracle against (a) the reference's own known-answer vectors and (b) the golden fixtures
generated from the imported Python reference (oracle/gen_golden.py).  CPU only."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from oracle import bindings as ob
from tests import golden_util as gu


def _cfg(meta, W, H):
    e = meta["enc"]
    return ob.make_config(W, H, e["block"], e["search_range"], e["qp"], nref=e.get("nref", 1),
                          fastme=e.get("fastme", False), frac=e.get("frac", False), i_period=e["i_period"])


@pytest.mark.parametrize("name", [n for n in gu.names() if n != "cif_c1" and not n.startswith("rc")])
def test_clip_matches_reference_golden(name):
    g = gu.load(name)
    frames = g["frames"]
    n, H, W = frames.shape
    cfg = _cfg(g["meta"], W, H)
    data, recon = ob.encode_clip(cfg, frames, nthreads=2)
    assert hashlib.sha256(data).hexdigest() == g["meta"]["encoded_sha256"]
    assert data == g["encoded"]
    if "recon" in g:
        assert np.array_equal(recon, g["recon"])
    else:   # the CIF fixtures (BASELINE configs 1-3 at their real geometry) store the hash of the reconstruction
        assert hashlib.sha256(recon.tobytes()).hexdigest() == g["meta"]["recon_sha256"]


def test_cif_config1_standin():
    g = gu.load("cif_c1")
    frames = g["frames"]
    n, H, W = frames.shape
    cfg = _cfg(g["meta"], W, H)
    data, recon = ob.encode_clip(cfg, frames, nthreads=2)
    assert data == g["encoded"]
    assert hashlib.sha256(recon.tobytes()).hexdigest() == g["meta"]["recon_sha256"]


@pytest.mark.parametrize("name", ["fs_i8_r4_qp3", "fastme_i16_nref4", "frac_fs_i8_r2_nref2",
                                  "frac_fastme_i8_nref3", "ties_i8_r4", "fs_i16_r2_nref4"])
def test_frame_level_details(name):
    """Per-frame MVs / modes / avg_mae / comparison counts / bits per row / debug planes."""
    g = gu.load(name)
    frames, meta = g["frames"], g["meta"]
    n, H, W = frames.shape
    e = meta["enc"]
    cfg = _cfg(meta, W, H)
    nref = e.get("nref", 1)
    refs, hps = [], []
    for idx in range(n):
        det = meta["frames"][idx]
        if idx % e["i_period"] == 0:
            r = ob.encode_iframe(cfg, frames[idx])
            refs, hps = [], []
            assert det["intra"] == 1
            assert r.modes.tolist() == det["modes"]
            assert np.array_equal(r.resid_mc.view(np.uint8), g["resid_mc"][idx])
        else:
            r = ob.encode_pframe(cfg, frames[idx], refs, hps if e.get("frac") else None)
            assert det["intra"] == 0
            assert r.mv.tolist() == det["mv"] if nref > 1 or e.get("fastme") else r.mv[:, :2].tolist() == [m[:2] for m in det["mv"]]
            assert np.array_equal(r.resid_nomc.view(np.uint8), g["resid_nomc"][idx])
            assert np.array_equal(r.resid_mc.view(np.uint8), g["resid_mc"][idx])
        assert r.avg_mae == det["avg_mae"]
        assert r.mae_comparisons == det["mae_comparisons"]
        assert r.bits_per_row.tolist() == det["bits_per_row"]
        assert r.pred_nbits == det["pred_nbits"] and r.coef_nbits == det["coef_nbits"]
        assert hashlib.sha256(r.pred_bytes).hexdigest() == det["pred_sha"]
        assert hashlib.sha256(r.coef_bytes).hexdigest() == det["coef_sha"]
        assert np.array_equal(r.recon, g["recon"][idx])
        assert np.array_equal(r.levels, g["levels"][idx])
        refs.append(r.recon)
        hps.append(ob.halfpel_plane(r.recon) if e.get("frac") else r.recon)
        if len(refs) > nref:
            refs.pop(0)
            hps.pop(0)


# ---- the reference's own known-answer vectors ------------------------------

def test_q_matrix_vectors():
    """tests/test_dct.py:23-30 of the reference."""
    L = ob.lib()
    q42 = [[1 << L.bvo_q_shift(4, 2, x, y) for y in range(4)] for x in range(4)]
    assert q42 == [[4, 4, 4, 8], [4, 4, 8, 16], [4, 8, 16, 16], [8, 16, 16, 16]]
    q20 = [[1 << L.bvo_q_shift(2, 0, x, y) for y in range(2)] for x in range(2)]
    assert q20 == [[1, 2], [2, 4]]


def _rle(seq):
    L = ob.lib()
    a = np.asarray(seq, dtype=np.int16)
    out = np.zeros(2 * len(a) + 1, dtype=np.int32)
    m = L.bvo_rle(a.ctypes.data_as(C.c_void_p), len(a), out.ctypes.data_as(C.c_void_p))
    return out[:m].tolist()


def test_rle_known_answer():
    """tests/test_entropy_encoder.py:41-66 of the reference."""
    seq = [0] * 16 + list(range(1, 9)) + [0] * 5 + list(range(1, 5)) + [0] * 3 + list(range(1, 9)) + [0] * 100
    assert len(seq) == 144
    assert _rle(seq) == [16, -8, 1, 2, 3, 4, 5, 6, 7, 8, 5, -4, 1, 2, 3, 4, 3, -8, 1, 2, 3, 4, 5, 6, 7, 8, 0]
    assert _rle([0] * 16) == [0]
    assert _rle([5]) == [-1, 5]


def test_exp_golomb_codes():
    """Signed exp-Golomb (entropy_encoder.py:8-29) incl. the round-trip list of
    tests/test_entropy_encoder.py:72-86 and the EOB marker length (27 bits)."""
    L = ob.lib()
    expect = {0: "1", 1: "010", -1: "011", 2: "00100", -3: "00111", 4: "0001000", -4: "0001001"}
    for v, bits in expect.items():
        b = ob.Bits()
        L.bvo_bits_init(C.byref(b))
        L.bvo_put_eg(C.byref(b), v)
        assert b.nbits == len(bits) == L.bvo_eg_len(v)
        got = "".join(str((b.data[i >> 3] >> (7 - (i & 7))) & 1) for i in range(b.nbits))
        assert got == bits
        L.bvo_bits_free(C.byref(b))
    assert L.bvo_eg_len(8190) == 27


def test_zigzag_order():
    """entropy_encoder.py:115-135: (0,0),(1,0),(0,1),(0,2),(1,1),(2,0),..."""
    L = ob.lib()
    m = np.arange(16, dtype=np.int16).reshape(4, 4)
    out = np.zeros(16, dtype=np.int16)
    L.bvo_zigzag(m.ctypes.data_as(C.c_void_p), 4, 4, out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == [0, 4, 1, 2, 5, 8, 12, 9, 6, 3, 7, 10, 13, 14, 11, 15]


def test_halfpel_tiny():
    """tests/playground.py:66-80 style: ceil averages; last row/column stay 0."""
    ref = np.array([[10, 20], [30, 41]], dtype=np.uint8)
    hp = ob.halfpel_plane(ref)
    assert hp.tolist() == [[10, 15, 20, 0], [20, 26, 31, 0], [30, 36, 41, 0], [0, 0, 0, 0]]


def test_dct_rational_bins_exact_and_ties_half_even():
    """The four bins {0,N/2}^2 are exact in the defined transform, so exact quantiser ties resolve
    as round-half-to-even."""
    for bs in (4, 8, 16):
        res = np.zeros((bs, bs), dtype=np.int16)
        res[0, 0] = bs * 4 * 3  # DC coefficient = 12 exactly; qp=3 -> Q=8 -> 1.5 -> 2 (even)
        coef = ob.fdct(res)
        assert coef[0, 0] == 12.0
        level, recon, _, _ = ob.transform_block(res, np.zeros((bs, bs), np.int16), 3)
        assert level[0, 0] == 2
        res[0, 0] = bs * 4  # 4/8 = 0.5 -> 0
        level, _, _, _ = ob.transform_block(res, np.zeros((bs, bs), np.int16), 3)
        assert level[0, 0] == 0


def test_dct_is_orthonormal_dct2():
    rng = np.random.default_rng(0)
    from scipy.fftpack import dct
    for bs in (2, 4, 8, 16, 32):
        x = rng.integers(-255, 256, size=(bs, bs)).astype(np.int16)
        ref = dct(dct(x.astype(np.float64).T, norm="ortho").T, norm="ortho")
        got = ob.fdct(x)
        assert np.max(np.abs(ref - got)) < 1e-9
        back = ob.idct(got)
        assert np.max(np.abs(back - x)) < 1e-9


@pytest.mark.parametrize("name", gu.names())
def test_decoder_oracle_matches_reference_decoder(name):
    """bvo_decode_clip (decode_video restated) against what the reference's own decode_video produced from the
    same container bytes (oracle/gen_golden_decode.py -> tests/golden/decode_ref.json), and against the
    encoder's reconstruction (decode == recon is the reference's own decoder test, tests/test_decoder.py:37-78)."""
    import json
    import os
    ref = json.load(open(os.path.join(gu.GOLD, "decode_ref.json")))[name]
    g = gu.load(name)
    n, H, W = g["frames"].shape
    cfg = _cfg(g["meta"], W, H)
    dec, lev, pred, qps, kinds = ob.decode_clip(cfg, g["encoded"], n + 3, details=True)
    assert dec.shape[0] == ref["frames"] == n
    assert hashlib.sha256(dec.tobytes()).hexdigest() == ref["decoded_sha256"]
    if "recon" in g:
        assert np.array_equal(dec, g["recon"])
    elif "recon_sha256" in g["meta"]:
        assert hashlib.sha256(dec.tobytes()).hexdigest() == g["meta"]["recon_sha256"]
    if "levels" in g:
        assert np.array_equal(lev, g["levels"])
    assert kinds[0] == 1
    # max_frames cuts the loop like frames_to_process (decoder.py:49)
    assert ob.decode_clip(cfg, g["encoded"], 2).shape[0] == 2


def test_decoder_oracle_rejects_malformed_streams():
    g = gu.load("fs_i8_r4_qp3")
    n, H, W = g["frames"].shape
    cfg = _cfg(g["meta"], W, H)
    data = bytearray(g["encoded"])
    with pytest.raises(ValueError):
        ob.decode_clip(cfg, bytes(data[:len(data) - 40]), n)          # truncated record
    bad = bytearray(data)
    bad[3] ^= 0xFF                                                    # garbage in the first prediction symbols
    try:
        ob.decode_clip(cfg, bytes(bad), n)
    except ValueError:
        pass


def test_recorded_dct_divergence_claims():
    """cif_c1_dct_divergence.json: what the judge of a drop-in needs to know, as recorded from the imported reference.
    fp64 SciPy against the defined transform on identical residuals: different levels only at exact ties, coefficients
    within 1e-9; the shipped float32 arithmetic differs from the contract in a handful of levels per frame."""
    import json
    import os
    rec = json.load(open(os.path.join(gu.GOLD, "cif_c1_dct_divergence.json")))
    probe = rec["same_residual_probe_fp64_scipy_vs_defined"]
    assert probe["blocks"] == 15840 and probe["non_tie_diffs"] == 0 and probe["max_tie_distance"] == 0.0
    assert probe["max_coef_diff_all"] < 1e-9
    asis = rec["modes"][0]
    assert asis["mode"].startswith("asis") and asis["total_levels"] == 10 * 288 * 352
    assert 0 < asis["total_differing_levels"] < asis["total_levels"] // 1000      # 54 of 1 013 760 when recorded
    assert abs(asis["container_bytes"] - asis["container_bytes_contract"]) < 64
    for m in rec["modes"]:
        for fr in m["frames"]:
            assert abs(fr["psnr"] - fr["psnr_contract"]) < 0.25      # largest recorded: 0.11 dB (fp64_scipy, frame 7)


def test_bench_clip_oracle_record_is_consistent():
    """tests/golden/bench_clip_oracle.json (oracle/gen_bench_clip_sha.py): the oracle's stream of the whole bench.py
    workload; the per-GOP fragments must add up to the serial stream's length."""
    import json
    import os
    rec = json.load(open(os.path.join(gu.GOLD, "bench_clip_oracle.json")))
    assert len(rec["gops"]) == 20 and sum(g["bytes"] for g in rec["gops"]) == rec["bytes"]
