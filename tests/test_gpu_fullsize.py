"""BASELINE.json's full sizes: oracle comparisons where the CPU finishes in seconds, size-independent properties
(decode(encode(x)) == reconstruction, GOP streams concatenate, spot-checked motion vectors) where it does not."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def test_1080p_r32_gop_matches_oracle_and_roundtrips():
    """configs[3]: 1920x1088, i=16, r=32 full search.  One I + two P frames against the oracle (bit-exact stream and
    reconstruction), then 2 GOPs x 3 frames through GOP lanes and lane groups: decode == reconstruction."""
    import basic_video_codec_b200 as bvc
    from oracle import bindings as ob
    W, H, bs, r, qp = 1920, 1088, 16, 32, 4
    frames = synth.moving_clip(1080, H, W, 6, step=6, clamp=96, noise=2)
    cfg = ob.make_config(W, H, bs, r, qp, nref=1, i_period=3)
    want, want_recon = ob.encode_clip(cfg, frames[:3], nthreads=1)
    with bvc.Context(W, H, bs, r, qp, 1, False, False, 3, device=0, max_lanes=2) as ctx:
        data, recon = ctx.encode_clip(frames, want_recon=True)
        assert data[:len(want)] == want
        assert np.array_equal(recon[:3], want_recon)
        assert np.array_equal(ctx.decode_clip(data, 6), recon)
        ctx.set_lane_groups(1)
        assert ctx.encode_clip(frames)[0] == data


def test_4k_r64_4refs_spot_checks_and_roundtrip():
    """configs[4]: 3840x2160, i=16, r=64, 4 references.  One GOP of 4 frames: motion vectors / SADs of sampled blocks
    (corners, edges, interior) of the 3-reference P frame against the oracle's full search; decode == reconstruction."""
    import basic_video_codec_b200 as bvc
    from oracle import bindings as ob
    W, H, bs, r, qp = 3840, 2160, 16, 64, 4
    frames = synth.moving_clip(2160, H, W, 4, step=3, clamp=48, noise=2)
    with bvc.Context(W, H, bs, r, qp, 4, False, False, 8, device=0, max_lanes=1) as ctx:
        data, recon = ctx.encode_clip(frames, want_recon=True)
        assert np.array_equal(ctx.decode_clip(data, 4), recon)
        mv, sad, _ = ctx.me_search(frames[3], [recon[0], recon[1], recon[2]])
    bw, bh = W // bs, H // bs
    rng = np.random.default_rng(5)
    picks = [(0, 0), (bw - 1, 0), (0, bh - 1), (bw - 1, bh - 1), (bw // 2, 0), (0, bh // 2)] + \
            [(int(rng.integers(bw)), int(rng.integers(bh))) for _ in range(10)]
    for bx, by in picks:
        omv, osad = ob.full_search_block(frames[3], [recon[0], recon[1], recon[2]], bx * bs, by * bs, bs, r)
        b = by * bw + bx
        assert mv[b].tolist() == list(omv) and int(sad[b]) == osad, (bx, by)


@pytest.mark.parametrize("W,H,bs,r,nref,fastme,frac", [
    (7680, 4320, 16, 16, 1, False, False),    # 8K: 129 600 blocks per frame, 270 x 480 block grid
    (7680, 4320, 16, 16, 2, True, False),     # 8K FastME: chains of 129 600 blocks (transfer tables: 2 lanes in flight)
    (4096, 2176, 8, 8, 2, False, True),       # half-pel planes of a 4K-wide frame, 8x8 blocks
    (1936, 1096, 8, 4, 1, False, False),      # width and height that are multiples of 8 but not of 16 (pitch padding)
])
def test_large_and_odd_geometries_roundtrip(W, H, bs, r, nref, fastme, frac):
    """Sizes beyond BASELINE's: the oracle would take minutes, so the size-independent property is checked --
    decode(encode(x)) equals the encoder's reconstruction -- plus oracle motion vectors of a few sampled blocks."""
    import basic_video_codec_b200 as bvc
    from oracle import bindings as ob
    n, ip = 4, 2
    base = synth.moving_clip(5, 544, 960, n, step=3, clamp=24)
    frames = np.ascontiguousarray(np.tile(base, (1, (H + 543) // 544, (W + 959) // 960))[:, :H, :W])
    with bvc.Context(W, H, bs, r, 4, nref, fastme, frac, ip, device=0, max_lanes=2) as ctx:
        data, recon = ctx.encode_clip(frames, want_recon=True)
        assert np.array_equal(ctx.decode_clip(data, n), recon)
        if not fastme and not frac:
            mv, sad, _ = ctx.me_search(frames[1], [recon[0]])
            bw, bh = W // bs, H // bs
            for bx, by in [(0, 0), (bw - 1, bh - 1), (bw // 2, bh // 3), (bw - 1, 0)]:
                omv, osad = ob.full_search_block(frames[1], [recon[0]], bx * bs, by * bs, bs, r)
                b = by * bw + bx
                assert mv[b].tolist() == list(omv) and int(sad[b]) == osad, (bx, by)


def test_1080p_whole_gops_through_20_lanes_match_the_oracle_record():
    """configs[3] as bench.py runs it: 20 GOP lanes, two lane groups, whole 30-frame GOPs.  Lanes alternate between GOP 0 and
    GOP 1 of the bench clip (the first 60 frames of the seeded generator); every lane's container fragment must have the
    sha256 the CPU oracle produced for that GOP (tests/golden/bench_clip_oracle.json, 258 s of 8 host cores for the whole
    600-frame clip; bench.py checks the whole stream's hash on every run)."""
    import hashlib
    import json
    import os
    import basic_video_codec_b200 as bvc
    from basic_video_codec_b200.sharding import split_container_by_gop
    from tests import golden_util as gu
    rec = json.load(open(os.path.join(gu.GOLD, "bench_clip_oracle.json")))
    W, H, bs, r, qp, ip = 1920, 1088, 16, 32, 4, 30
    two = synth.moving_clip(1080, H, W, 2 * ip, step=6, clamp=96, noise=2)
    frames = np.ascontiguousarray(np.tile(two, (10, 1, 1)))
    with bvc.Context(W, H, bs, r, qp, 1, False, False, ip, device=0, max_lanes=20) as ctx:
        assert ctx.lane_groups == 2
        data, _ = ctx.encode_clip(frames)
    parts = split_container_by_gop(data, [ip] * 20)
    for lane, p in enumerate(parts):
        want = rec["gops"][lane % 2]
        assert len(p) == want["bytes"] and hashlib.sha256(p).hexdigest() == want["sha256"], f"lane {lane}"
