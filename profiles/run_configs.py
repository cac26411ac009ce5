#!/usr/bin/env python3
"""Secondary workloads of BASELINE.json (configs 0, 1, 2, 4) on one B200: throughput through the public
clip API (host buffers), per-kernel device time, and a parity spot check of the first GOP against the CPU
oracle (GOPs are independent, so the first GOP of the stream must equal the oracle's encoding of it).
Foreman is not available (git-LFS pointer), so the CIF configs use the synthetic stand-in.
Usage: python profiles/run_configs.py [name ...]   -> one JSON line per workload."""
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import basic_video_codec_b200 as bvc  # noqa: E402
from basic_video_codec_b200.sharding import split_container_by_gop  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from tests import synth  # noqa: E402

WORKLOADS = {
    # name: (W, H, frames, block, r, qp, I_Period, nref, fastme, frac, lanes, synth kwargs)
    "c0_cif_i8_r4": (352, 288, 10, 8, 4, 3, 8, 1, False, False, 2, dict(step=3, clamp=16)),
    "c1_cif_fastme_nref4": (352, 288, 300, 16, 16, 3, 8, 4, True, False, 38, dict(step=2, clamp=24)),
    "c2_cif_halfpel_r4": (352, 288, 296, 16, 4, 3, 8, 1, False, True, 37, dict(step=2, clamp=16)),
    "c4_4k_r64_nref4": (3840, 2160, 512, 16, 64, 4, 8, 4, False, False, 32, dict(step=3, clamp=48)),
    "c3_1080p_r32": (1920, 1088, 600, 16, 32, 4, 30, 1, False, False, 20, dict(step=6, clamp=96)),
    # not BASELINE configurations: FastME at the headline geometry (a chain of 8160 blocks per frame) and with half-pel
    "x_1080p_fastme_nref4": (1920, 1088, 600, 16, 16, 4, 30, 4, True, False, 20, dict(step=6, clamp=96)),
    "x_cif_halfpel_fastme_nref2": (352, 288, 296, 16, 4, 3, 8, 2, True, True, 37, dict(step=2, clamp=16)),
    # round 2: the narrow-range search kernel on a GPU-filling size (config 3's search at 1080p), the reference's own ranges
    "x_1080p_halfpel_r4": (1920, 1088, 600, 16, 4, 4, 30, 1, False, True, 20, dict(step=3, clamp=48)),
    "x_1080p_i16_r4": (1920, 1088, 600, 16, 4, 4, 30, 1, False, False, 20, dict(step=3, clamp=48)),
    "x_1080p_i8_r2": (1920, 1088, 600, 8, 2, 3, 30, 1, False, False, 20, dict(step=2, clamp=48)),
    "x_1080p_i8_r3": (1920, 1088, 600, 8, 3, 3, 30, 1, False, False, 20, dict(step=2, clamp=48)),
    "x_cif_i16_r2_nref4": (352, 288, 296, 16, 2, 3, 8, 4, False, False, 37, dict(step=2, clamp=16)),
}
FASTME_MODE = int(os.environ.get("BVC_FASTME_MODE", "0"))   # bvc_set_fastme_direct: 0 auto, 1 direct, 2 serial map walk, 3 window walk, 4 tables


def run(name):
    W, H, n, bs, r, qp, ip, nref, fastme, frac, lanes, sk = WORKLOADS[name]
    t0 = time.time()
    frames = synth.moving_clip(zlib.crc32(name.encode()) % 1000 + 7, H, W, n, **sk)   # same content in every process
    tgen = time.time() - t0
    out = np.empty(n * W * H // 2 + (1 << 20), np.uint8)
    # page-locked host buffers, as in bench.py: with small search ranges the clip is bound by its upload, and pageable memory
    # (the round's first run of these lines) halves-to-quarters the transfer rate
    from basic_video_codec_b200._lib import host_register, host_unregister
    host_register(frames)
    host_register(out)
    with bvc.Context(W, H, bs, r, qp, nref, fastme, frac, ip, device=0, max_lanes=lanes) as ctx:
        if fastme:
            ctx.set_fastme_direct(FASTME_MODE)
        ctx.encode_clip_into(frames, out)                      # warm-up
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            ln = ctx.encode_clip_into(frames, out)
        dt = (time.perf_counter() - t0) / reps
        _, clip_ms = ctx.last_kernel_times()
        ctx.set_lane_groups(1)                                  # per-kernel times need the kernels serialised on one stream
        ctx.encode_clip_into(frames, out)
        kt, _ = ctx.last_kernel_times()
        ctx.set_lane_groups(2)
        work = ctx.me_work_per_frame(1)
        # the same clip resident in HBM (what the kernels alone sustain)
        ctx.clip_upload(frames)
        ctx.encode_clip_resident(n, out)
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.encode_clip_resident(n, out)
        dt_res = (time.perf_counter() - t0) / reps
    host_unregister(frames)
    host_unregister(out)
    data = out[:ln].tobytes()
    # parity spot check: first GOP (and the last, possibly short, one) against the oracle
    gops = [(f0, min(ip, n - f0)) for f0 in range(0, n, ip)]
    parts = split_container_by_gop(data, [g[1] for g in gops])
    cfg = ob.make_config(W, H, bs, r, qp, nref=nref, fastme=fastme, frac=frac, i_period=ip)
    checked = []
    t0 = time.perf_counter()
    for gi in sorted({0, len(gops) - 1}) if W <= 1920 else [0]:
        f0, nf = gops[gi]
        nchk = nf if W <= 352 else min(nf, 3)                  # oracle time at 1080p/4K: a few seconds per P frame
        want, _ = ob.encode_clip(cfg, frames[f0:f0 + nchk], want_recon=False)
        got = parts[gi]
        if nchk < nf:   # compare the first nchk frame records only
            got = b"".join(split_container_by_gop(got, [1] * nf)[:nchk])
        checked.append(bool(got == want))
    t_oracle = time.perf_counter() - t0
    line = {"workload": name, "geometry": f"{W}x{H} i={bs} r={r} qp={qp} I_Period={ip} nRef={nref} fastME={fastme} frac={frac}",
            "frames": n, "lanes": lanes, "e2e_frames_per_s": n / dt, "ms_per_clip": dt * 1e3, "device_ms": clip_ms,
            "resident_frames_per_s": n / dt_res, "resident_ms_per_clip": dt_res * 1e3, "host_buffers": "page-locked",
            "kernel_ms": {k: v[0] for k, v in kt.items()}, "kernel_launches": {k: v[1] for k, v in kt.items()},
            "bitstream_bytes": ln, "oracle_gops_checked": checked, "oracle_check_s": t_oracle, "synth_s": tgen}
    if fastme:
        line["fastme_mode"] = FASTME_MODE
    if not fastme and kt["me"][0] > 0:
        line["me_px_absdiff_per_frame_1ref"] = work
        # frame k of a GOP sees min(k, nRef) references (deque cleared at the I frame)
        tot = sum(work * min(k, nref) for f0, nf in gops for k in range(1, nf))
        line["me_tpx_per_s"] = tot / (kt["me"][0] * 1e-3) / 1e12
    print(json.dumps(line), flush=True)
    assert all(checked), f"{name}: GPU stream differs from the oracle"


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["c0_cif_i8_r4", "c1_cif_fastme_nref4", "c2_cif_halfpel_r4"]):
        run(nm)
