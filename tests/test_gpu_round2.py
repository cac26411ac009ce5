"""Round-2 GPU tests: hazards the round-1 review named (stale resident clip, wavefront grids beyond resident
capacity), the generalised tiled search, the sharded job and rate control on the clip path."""
import os

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _ob():
    from oracle import bindings as ob
    return ob


def _ctx(W, H, bs, r, qp, nref=1, fastme=False, frac=False, ip=1, lanes=1):
    import basic_video_codec_b200 as bvc
    return bvc.Context(W, H, bs, r, qp, nref, fastme, frac, ip, device=0, max_lanes=lanes)


def test_resident_clip_is_invalidated_by_frame_calls():
    """bvc_clip_upload -> frame-level call (overwrites plane 0 of the input pool) -> bvc_encode_clip_resident must refuse
    instead of silently encoding the wrong frames; the same after a host-buffer clip call."""
    W, H, bs, r, qp = 96, 64, 16, 8, 3
    clip = synth.moving_clip(31, H, W, 6, step=3, clamp=16)
    other = synth.moving_clip(32, H, W, 6, step=3, clamp=16)
    with _ctx(W, H, bs, r, qp, ip=3, lanes=2) as ctx:
        ctx.clip_upload(clip)
        out, n = ctx.encode_clip_resident(6)
        want = bytes(out[:n])
        assert ctx.encode_clip(clip)[0] == want
        # host path overwrote the pool (with `clip` again, but the library cannot know that)
        with pytest.raises(ValueError):
            ctx.encode_clip_resident(6)
        ctx.clip_upload(clip)
        ctx.encode_pframe(other[1], [other[0]])
        with pytest.raises(ValueError):
            ctx.encode_clip_resident(6)
        ctx.clip_upload(clip)
        out, n = ctx.encode_clip_resident(6)
        assert bytes(out[:n]) == want


def test_iframe_wavefront_grid_beyond_resident_capacity():
    """32 I frames of 64 x 6000 with 4x4 blocks: 1500 block rows x 4 lane groups of the wavefront = 6000 one-warp CTAs,
    more than the 148 x 32 the GPU can hold at once.  Rows are claimed by start-order tickets, so a waiting row's
    producer is always resident; the streams and reconstructions must equal the oracle's, and the decoder's wavefront
    (same scheme) must give the reconstruction back."""
    ob = _ob()
    W, H, bs, qp, n = 64, 6000, 4, 2, 32
    base = synth.moving_clip(77, 600, W, n, step=2, clamp=8)
    frames = np.ascontiguousarray(np.tile(base, (1, 10, 1)))
    cfg = ob.make_config(W, H, bs, 1, qp, nref=1, i_period=1)
    want, want_recon = ob.encode_clip(cfg, frames)
    with _ctx(W, H, bs, 1, qp, ip=1, lanes=n) as ctx:
        for groups in (1, 2):
            ctx.set_lane_groups(groups)
            data, recon = ctx.encode_clip(frames, want_recon=True)
            assert data == want
            assert np.array_equal(recon, want_recon)
        assert np.array_equal(ctx.decode_clip(data, n), recon)


# ---- generalised search kernels: every (i, r), integer and half-pel ------------------------------------------------------
# narrow kernel (2r < i), tiled kernel with a padded vertical range (2r >= i, 2r % i != 0), tiled exact (2r % i == 0)
GEN_ME_CASES = [(bs, r, frac) for bs in (4, 8, 16) for r in (1, 2, 3, 5, 7) for frac in (False, True)] + \
               [(16, 4, False), (16, 4, True), (16, 6, False), (16, 12, False), (16, 12, True), (8, 9, False), (8, 6, True), (4, 9, False)]


@pytest.mark.parametrize("bs,r,frac", GEN_ME_CASES)
def test_generalised_search_matches_oracle(bs, r, frac):
    """Frames of 176 x 144 (QCIF: 1.4 tiles of the narrow kernel side by side, 2.25 stacked at i=16), two references,
    moving and tie-heavy content: motion vectors, reference indices and SADs of every block against the oracle."""
    ob = _ob()
    W, H, nref = 176, 144, 2
    for content in ("moving", "poster"):
        clip = (synth.moving_clip(300 + bs + r, H, W, nref + 1, step=min(r, 5), clamp=16) if content == "moving"
                else synth.posterised_clip(400 + bs + r, H, W, nref + 1))
        cur, refs = clip[-1], [clip[i] for i in range(nref)]
        cfg = ob.make_config(W, H, bs, r, 3, nref=nref, frac=frac)
        planes = [ob.halfpel_plane(x) for x in refs] if frac else refs
        mv_o, sad_o, cmp_o = ob.me_frame(cfg, cur, planes)
        with _ctx(W, H, bs, r, 3, nref, False, frac) as ctx:
            mv_g, sad_g, cmp_g = ctx.me_search(cur, refs)
        assert np.array_equal(sad_g, sad_o), content
        assert np.array_equal(mv_g, mv_o), content
        assert cmp_g == cmp_o


def test_config3_geometry_cif_halfpel_r4_clip_matches_oracle():
    """BASELINE configs[2] geometry (CIF, i=16, r=4, half-pel) through the clip call: 2 GOPs of 4 frames in lanes."""
    ob = _ob()
    W, H, bs, r, qp = 352, 288, 16, 4, 4
    frames = synth.moving_clip(352, H, W, 8, step=3, clamp=16)
    cfg = ob.make_config(W, H, bs, r, qp, nref=2, frac=True, i_period=4)
    want, want_recon = ob.encode_clip(cfg, frames)
    with _ctx(W, H, bs, r, qp, 2, False, True, ip=4, lanes=2) as ctx:
        data, recon = ctx.encode_clip(frames, want_recon=True)
    assert np.array_equal(recon, want_recon)
    assert data == want


def test_tail_split_many_lanes_matches_oracle(monkeypatch):
    """110 GOP lanes of 128 x 192 frames (i=16, r=16): 660 search tiles per step, more than the GPU holds at once (592 at
    four 160-thread CTAs per SM), so the tiles of the last wave run as one-row CTAs.  Stream equal to the oracle's and
    to the unsplit launch."""
    ob = _ob()
    W, H, bs, r, qp, ip, ngop = 128, 192, 16, 16, 4, 2, 110
    base = synth.moving_clip(61, H, W, 8, step=5, clamp=24)
    frames = np.ascontiguousarray(np.concatenate([np.roll(base[(2 * g) % 6: (2 * g) % 6 + 2], g, axis=2) for g in range(ngop)]))
    cfg = ob.make_config(W, H, bs, r, qp, nref=1, i_period=ip)
    want, _ = ob.encode_clip(cfg, frames, want_recon=False)
    with _ctx(W, H, bs, r, qp, ip=ip, lanes=ngop) as ctx:
        ctx.set_lane_groups(1)
        assert ctx.encode_clip(frames)[0] == want
    monkeypatch.setenv("BVC_TAIL_SPLIT", "0")
    with _ctx(W, H, bs, r, qp, ip=ip, lanes=ngop) as ctx:
        ctx.set_lane_groups(1)
        assert ctx.encode_clip(frames)[0] == want


@pytest.mark.parametrize("tall,split", [("1", "1"), ("1", "0"), ("-1", "1")])
def test_tall_search_tiles_match_oracle(tall, split, monkeypatch):
    """i=16, r=32: the headline search shape and its tall variant (eight stacked block rows per CTA, taken by large
    launches -- forced here with BVC_ME_TALL) on a frame whose 13 x 11 blocks fill neither the last tile column nor the
    last tile row, 75 GOP lanes in one launch (300 tall tiles: one whole wave of 296 plus a split tail): stream and
    reconstruction equal to the oracle's, every lane through the extra jobs of the last candidate column."""
    ob = _ob()
    W, H, bs, r, qp, ip, ngop = 208, 176, 16, 32, 3, 2, 75
    base = synth.moving_clip(67, H, W, 8, step=7, clamp=40)
    frames = np.ascontiguousarray(np.concatenate([np.roll(base[(2 * g) % 6: (2 * g) % 6 + 2], 3 * g, axis=2) for g in range(ngop)]))
    cfg = ob.make_config(W, H, bs, r, qp, nref=1, i_period=ip)
    want, want_recon = ob.encode_clip(cfg, frames)
    monkeypatch.setenv("BVC_ME_TALL", tall)
    monkeypatch.setenv("BVC_TAIL_SPLIT", split)
    with _ctx(W, H, bs, r, qp, ip=ip, lanes=ngop) as ctx:
        ctx.set_lane_groups(1)
        data, recon = ctx.encode_clip(frames, want_recon=True)
    assert data == want
    assert np.array_equal(recon, want_recon)


@pytest.mark.parametrize("bs", [8, 16])
def test_iframe_quad_and_single_warp_wavefronts_match_oracle(bs, monkeypatch):
    """The I-frame wavefront with four warps per block pair (few CTAs in flight) and with one (BVC_IQUAD=0, and any launch of
    more than two CTAs per SM) are the same arithmetic: streams, reconstructions and the frame-level outputs (modes, levels,
    residual plane) equal the oracle's for both, with an odd number of lanes (a half-filled warp) and the transform's
    P-frame path (looping over work units under a CTA cap) after them."""
    ob = _ob()
    W, H, r, qp, ip, n = 160, 96, 4, 3, 3, 9
    frames = synth.moving_clip(91 + bs, H, W, n, step=2, clamp=8)
    cfg = ob.make_config(W, H, bs, r, qp, nref=1, i_period=ip)
    want, want_recon = ob.encode_clip(cfg, frames)
    fo = ob.encode_iframe(cfg, frames[0])
    for quad, cap in (("1", "0"), ("0", "0"), ("1", "7")):
        monkeypatch.setenv("BVC_IQUAD", quad)
        monkeypatch.setenv("BVC_TQ_CTAS", cap)
        with _ctx(W, H, bs, r, qp, ip=ip, lanes=3) as ctx:
            for groups in (1, 2):
                ctx.set_lane_groups(groups)
                data, recon = ctx.encode_clip(frames, want_recon=True)
                assert data == want, (quad, cap, groups)
                assert np.array_equal(recon, want_recon)
            fg = ctx.encode_iframe(frames[0])
            for k in ("recon", "levels", "modes", "resid_mc", "coef_bytes", "pred_bytes"):
                assert np.array_equal(np.asarray(getattr(fg, k)), np.asarray(getattr(fo, k))), (quad, k)


# ---- sharded job on the GPU ---------------------------------------------------------------------------------------------
def test_sharded_encoder_single_rank_equals_clip_call():
    """world = 1 through the sharding API (container left on the device, fetched into the shared buffer) == the plain
    clip call == the oracle; a short last GOP and the resident path included."""
    import basic_video_codec_b200 as bvc
    from basic_video_codec_b200 import sharding
    ob = _ob()
    W, H, bs, r, qp, ip, n = 96, 64, 16, 8, 3, 4, 10
    frames = synth.moving_clip(91, H, W, n, step=3, clamp=16)
    want, _ = ob.encode_clip(ob.make_config(W, H, bs, r, qp, nref=2, i_period=ip), frames, want_recon=False)
    ec = bvc.EncoderConfig(bs, r, ip, qp, nRefFrames=2)
    assert sharding.encode_clip_distributed(frames, ec) == want
    with sharding.ShardedEncoder(ec, W, H, n) as enc:
        assert bytes(enc.encode(frames)) == want
        enc.upload(frames)
        assert bytes(enc.encode(resident=True)) == want
        assert bytes(enc.encode(load_gop=lambda f0, k: frames[f0:f0 + k])) == want


def _shard_worker(rank, world, port, q):
    import hashlib
    import torch
    import torch.distributed as dist
    import basic_video_codec_b200 as bvc
    from basic_video_codec_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    W, H, bs, r, qp, ip, n = 192, 128, 16, 16, 4, 3, 15
    frames = synth.moving_clip(92, H, W, n, step=4, clamp=24)
    ec = bvc.EncoderConfig(bs, r, ip, qp, nRefFrames=1)
    with sharding.ShardedEncoder(ec, W, H, n, rank=rank, world=world, device=rank) as enc:
        out = enc.encode(frames)
        h1 = hashlib.sha256(bytes(out)).hexdigest() if out is not None else None
        enc.upload(frames)
        out = enc.encode(resident=True)
        h2 = hashlib.sha256(bytes(out)).hexdigest() if out is not None else None
    q.put((rank, h1, h2))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_encoder_two_gpus_equals_single_gpu_stream():
    """5 GOPs over 2 GPUs (NCCL for the length exchange and the barrier only): the stream rank 0 returns equals the
    single-GPU stream and the oracle's."""
    import hashlib
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ob = _ob()
    W, H, bs, r, qp, ip, n = 192, 128, 16, 16, 4, 3, 15
    frames = synth.moving_clip(92, H, W, n, step=4, clamp=24)
    want, _ = ob.encode_clip(ob.make_config(W, H, bs, r, qp, nref=1, i_period=ip), frames, want_recon=False)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_shard_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    res = dict((rk, (a, b)) for rk, a, b in (q.get(timeout=300) for _ in range(2)))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    hw = hashlib.sha256(want).hexdigest()
    assert res[0] == (hw, hw)
    assert res[1] == (None, None)


# ---- streaming clip path: input ring, per-wave container fragments, bounded stream slots ---------------------------------
@pytest.mark.parametrize("frac,nref,lanes,groups", [(False, 1, 2, 2), (True, 2, 3, 2), (False, 2, 1, 1), (False, 1, 4, 1)])
def test_many_waves_stream_through_the_rings(frac, nref, lanes, groups):
    """41 frames, I_Period 3 -> 14 GOPs (a short last one) in waves of 1-4 lanes: 4-14 waves, 42 steps through the
    6-step input ring, a container fragment per wave.  Stream, reconstruction and the resident path against the oracle."""
    ob = _ob()
    W, H, bs, r, qp, ip, n = 96, 64, 16, 8, 3, 3, 41
    frames = synth.moving_clip(123, H, W, n, step=3, clamp=16)
    cfg = ob.make_config(W, H, bs, r, qp, nref=nref, frac=frac, i_period=ip)
    want, want_recon = ob.encode_clip(cfg, frames)
    with _ctx(W, H, bs, r, qp, nref, False, frac, ip=ip, lanes=lanes) as ctx:
        ctx.set_lane_groups(groups)
        data, recon = ctx.encode_clip(frames, want_recon=True)
        assert data == want
        assert np.array_equal(recon, want_recon)
        assert ctx.encode_clip(frames)[0] == want            # a second pass over the same rings
        ctx.clip_upload(frames)
        out, ln = ctx.encode_clip_resident(n)
        assert bytes(out[:ln]) == want
        assert ctx.encode_clip_device(frames) == len(want)    # fragments appended on the device
        got = np.empty(len(want), np.uint8)
        ctx.container_download(got)
        assert bytes(got) == want


def test_small_output_buffer_and_small_stream_slots_are_reported_not_overrun():
    """Noise at QP 0: far more than the default reservations.  A too small host buffer and a too small device slot both
    come back as BVC_ERR_NOMEM (MemoryError); the binding's retry path ends with the oracle's stream."""
    ob = _ob()
    W, H, bs, r, qp, ip, n = 96, 64, 16, 4, 0, 2, 6
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, size=(n, H, W), dtype=np.uint8)
    want, _ = ob.encode_clip(ob.make_config(W, H, bs, r, qp, nref=1, i_period=ip), frames, want_recon=False)
    with _ctx(W, H, bs, r, qp, ip=ip, lanes=2) as ctx:
        small = np.empty(len(want) // 2, np.uint8)
        with pytest.raises(MemoryError):
            ctx.encode_clip_into(frames, small)
        ctx.set_stream_slot_bytes(2048)
        big = np.empty(2 * len(want), np.uint8)
        with pytest.raises(MemoryError, match="device slot"):
            ctx.encode_clip_into(frames, big)
        assert ctx.encode_clip(frames)[0] == want             # retries with the worst-case slot
        ctx.set_stream_slot_bytes(0)
        assert ctx.encode_clip(frames)[0] == want


def test_unbalanced_lane_groups_overflow_their_staging_region_and_the_call_is_repeated():
    """The container of a wave is staged in one region per lane group, sized from an estimate.  Two noisy GOPs in the first
    group and two flat ones in the second: the first part outgrows its region, the library enlarges the staging and asks for
    the call to be repeated (BVC_ERR_NOMEM, "staging"); the binding does that and ends with the oracle's stream -- through
    the host-buffer call, the resident call and the device-resident container of the sharded path."""
    ob = _ob()
    W, H, bs, r, qp, ip = 640, 480, 16, 2, 0, 6
    rng = np.random.default_rng(11)
    noisy = rng.integers(0, 256, size=(2 * ip, H, W), dtype=np.uint8)
    flat = np.full((2 * ip, H, W), 128, np.uint8)
    frames = np.ascontiguousarray(np.concatenate([noisy, flat]))
    n = frames.shape[0]
    want, _ = ob.encode_clip(ob.make_config(W, H, bs, r, qp, nref=1, i_period=ip), frames, want_recon=False, nthreads=8)
    assert len(want) > 2 * (n * W * H // 8 // 2 + (1 << 20))      # the first part alone is larger than two default regions
    with _ctx(W, H, bs, r, qp, ip=ip, lanes=4) as ctx:
        ctx.set_lane_groups(2)
        assert ctx.encode_clip(frames)[0] == want
    with _ctx(W, H, bs, r, qp, ip=ip, lanes=4) as ctx:
        ctx.set_lane_groups(2)
        ctx.set_stream_slot_bytes(W * H * 4)                       # 32 bits per pixel: only the staging is too small now
        out = np.empty(2 * len(want), np.uint8)
        ln = ctx.encode_clip_into(frames, out)
        assert bytes(out[:ln]) == want
    with _ctx(W, H, bs, r, qp, ip=ip, lanes=4) as ctx:
        ctx.set_lane_groups(2)
        ctx.set_stream_slot_bytes(W * H * 4)
        ctx.clip_upload(frames)
        out, ln = ctx.encode_clip_resident(n, np.empty(2 * len(want), np.uint8))
        assert bytes(out[:ln]) == want
        ln2 = ctx.encode_clip_device(None, n, cap_hint=2 * len(want))
        got = np.empty(ln2, np.uint8)
        ctx.container_download(got, 0, 0, ln2)
        assert bytes(got) == want
