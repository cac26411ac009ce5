"""GOP sharding host logic, incl. a world_size-2 gloo run on CPU.  The per-rank encoder is replaced by
the CPU oracle here (tests only); on the GPU box the same function drives libbvc_b200.so."""
import os
import socket

import numpy as np
import pytest

from tests import synth


def _oracle_encode(frames, ec, device):
    from oracle import bindings as ob
    n, H, W = frames.shape
    cfg = ob.make_config(W, H, ec.block_size, ec.search_range, ec.quantization_factor, nref=ec.nRefFrames,
                         fastme=ec.fastME, frac=ec.fracMeEnabled, i_period=ec.I_Period)
    return ob.encode_clip(cfg, frames, want_recon=False)[0]


def test_gop_maps():
    from basic_video_codec_b200 import sharding as sh
    assert sh.gop_ranges(10, 4) == [(0, 4), (4, 4), (8, 2)]
    assert sh.assign_gops(5, 2) == [[0, 1, 2], [3, 4]]      # contiguous runs: a rank's fragment is one slice of the stream
    assert sh.assign_gops(20, 8) == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11], [12, 13], [14, 15], [16, 17], [18, 19]]
    assert sh.assign_gops(2, 3) == [[0], [1], []]
    assert sh.scaling_ceiling(20, 8) == pytest.approx(20 / 3)   # 600 frames / I_Period 30 on 8 GPUs
    assert sh.scaling_ceiling(20, 4) == 4.0


def test_single_rank_equals_serial():
    from basic_video_codec_b200 import EncoderConfig, sharding as sh
    frames = synth.moving_clip(21, 32, 48, 10, step=2, clamp=8)
    ec = EncoderConfig(8, 2, 4, 3, nRefFrames=2)
    whole = _oracle_encode(frames, ec, 0)
    assert sh.encode_clip_distributed(frames, ec, encode_fn=_oracle_encode) == whole
    parts = sh.split_container_by_gop(whole, [4, 4, 2])
    assert b"".join(parts) == whole and len(parts) == 3


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from basic_video_codec_b200 import EncoderConfig, sharding as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = synth.moving_clip(22, 32, 48, 14, step=2, clamp=8)
    ec = EncoderConfig(8, 2, 4, 3, nRefFrames=1)
    touched = []

    def load(first, n):
        touched.append(first)
        return frames[first:first + n]

    out = sh.encode_clip_distributed(None, ec, rank=rank, world=world, encode_fn=_oracle_encode, load_gop=load, nframes=14, shape=(32, 48))
    q.put((rank, out, touched))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_concatenate_to_serial_stream():
    import torch.multiprocessing as mp
    from basic_video_codec_b200 import EncoderConfig
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, out, touched = q.get(timeout=120)
        res[rank] = (out, touched)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = synth.moving_clip(22, 32, 48, 14, step=2, clamp=8)
    whole = _oracle_encode(frames, EncoderConfig(8, 2, 4, 3, nRefFrames=1), 0)
    assert res[0][0] == whole          # rank 0 holds the serial stream
    assert res[1][0] is None
    assert res[0][1] == [0, 4] and res[1][1] == [8, 12]   # every rank only touched its own GOPs
