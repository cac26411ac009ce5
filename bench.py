#!/usr/bin/env python3
"""bench.py -- headline benchmark: encoded frames/s on synthetic 1920x1088 luma, 600 frames, i=16, r=32
integer full search, I_Period=30 (BASELINE.json configs[3]), GOP-sharded: one process per GPU, no data-path
collective (NCCL only for barrier / max / one length per rank).

  python bench.py [--gpus N] [--steps K] [--warmup W]              our CUDA path (libbvc_b200.so)
  python bench.py --impl reference [...]                           CPU arm: the oracle port on all host cores

One "step" = one pass of the encoder hot path over a 600-frame clip.
  value  : WEAK scaling -- every rank encodes its own 600-frame clip (N x 600 frames per step), clip resident in HBM
           (bit streams still come back to the host)
  e2e    : the same through the public API with the clip in pinned HOST memory (H2D inside the timed region)
  strong : the literal configs[3] -- ONE 600-frame clip, its 20 GOPs sharded over the N ranks through
           basic_video_codec_b200.sharding.ShardedEncoder (fragment lengths exchanged, every rank writes its fragment
           into a shared host buffer, rank 0 holds the serial stream); frames/s resident and e2e, the ceiling
           20 / ceil(20/N), and sha256(stream) checked against the single-GPU stream
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, BS, R, QP, IP, NFRAMES = 1920, 1088, 16, 32, 4, 30, 600
CLIP_SEED = 1080
# DRAM traffic per lane (one 1080p frame) of a launch, from the committed ncu --set full captures of the final build
# (profiles/r2_ncu_me_kernel.csv, profiles/r2_ncu_tq_kernel.csv: dram__bytes_read.sum + dram__bytes_write.sum of a 10-lane launch)
ME_TRAFFIC_BYTES_PER_LANE = (41.839104e6 + 0.392704e6) / 10       # 10-lane launch (two lane groups), profiles/r2_ncu_me_kernel.csv
TQ_TRAFFIC_BYTES_PER_LANE = (43.462656e6 + 6.845440e6) / 10       # profiles/r2_ncu_tq_kernel.csv
WORKLOAD = "synthetic 1920x1088 Y plane, 600 frames, i=16, r=32 full-search, I_Period=30, nRefFrames=1, QP=4 (BASELINE configs[3])"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def make_clip():
    """The workload's clip (one texture + seeded random walk + noise) in pinned host memory.  Every rank
    generates the same 600 frames: the weak-scaling ranks each encode their own copy, the strong-scaling ranks each take
    their GOPs from it, and all streams can be compared."""
    import torch
    from tests import synth
    buf = torch.empty((NFRAMES, H, W), dtype=torch.uint8, pin_memory=True)
    arr = buf.numpy()
    arr[:] = synth.moving_clip(CLIP_SEED, H, W, NFRAMES, step=6, clamp=96, noise=2)
    return arr, buf


# ---- CPU legs (the oracle is test infrastructure: it is only ever timed here, never on the product path) ----------------
def cpu_sample_clip(cores, band_h, gop_len):
    from tests import synth
    clip = synth.moving_clip(4242, H, W, gop_len, step=6, clamp=96, noise=2)
    gops = [np.roll(clip, shift=7 * g, axis=2)[:, :band_h, :] for g in range(cores)]
    return np.ascontiguousarray(np.concatenate(gops, axis=0))


def cpu_time(cores, band_h, gop_len, steps=1, warmup=0):
    """frames/s (scaled to full frames) of the oracle port: `cores` GOPs of I + (gop_len-1) P, one per host thread, on a band
    of band_h luma rows at full width."""
    from oracle import bindings as ob
    frames = cpu_sample_clip(cores, band_h, gop_len)
    cfg = ob.make_config(W, band_h, BS, R, QP, nref=1, i_period=gop_len)
    for _ in range(warmup):
        ob.encode_clip(cfg, frames, nthreads=cores, want_recon=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        ob.encode_clip(cfg, frames, nthreads=cores, want_recon=False)
    dt = time.perf_counter() - t0
    return frames.shape[0] * band_h / H * steps / dt, dt / steps, frames.shape[0]


SAMPLE_GOP, SAMPLE_BAND = 6, 416      # the bounded sample: I + 5 P on a band of 416 of the 1088 luma rows


def sample_note(cores, nfr, band_h, gop_len):
    return (f"{cores} GOPs x (I + {gop_len - 1} P) of the workload, one GOP per host thread, band of {band_h}/{H} luma rows at full "
            f"width ({nfr} frames per step, scaled by {band_h}/{H}); C restatement oracle/bvc_oracle.c (-O3, AVX2, OpenMP)")


def run_reference(args, rank):
    """CPU arm.  K timed steps of the bounded sample, plus ONE step of the real thing -- full-height frames, a whole
    I + 29 P GOP per host thread -- so the sample's extrapolation can be judged."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    val, step_s, nfr = cpu_time(cores, SAMPLE_BAND, SAMPLE_GOP, steps=args.steps, warmup=max(0, args.warmup))
    full = None
    if not args.skip_full:
        fval, fstep, fn = cpu_time(cores, H, IP, steps=1)
        full = {"value": fval, "unit": "frames/s", "seconds": fstep, "frames": fn,
                "what": f"{cores} whole GOPs (I + 29 P) at full height {W}x{H}, one per host thread, one step",
                "sample_over_full": val / fval}
    line = {
        "impl": "reference", "metric": "encoded frames/s", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8 SAD / int16 residual / f64 DCT", "data": "synthetic",
        "config": {"workload": WORKLOAD, "parallelism": f"{cores} host threads (OpenMP over GOPs)"},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": sample_note(cores, nfr, SAMPLE_BAND, SAMPLE_GOP)},
        "full_config_step": full,
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "value = the bounded sample (extrapolated to whole frames; errs in the CPU's favour, see full_config_step."
                "sample_over_full). C restatement of the reference's algorithm; the reference itself is pure Python/NumPy and "
                "~1000x slower: ~440 s per 1080p r=32 P frame on one core (BASELINE.md)",
    }
    print(json.dumps(line), flush=True)


def run_cpu_sample():
    """`--impl cpu-sample`: the cpu_baseline leg of our own line, run in a subprocess so that the GPU process never maps
    the oracle library."""
    cores = os.cpu_count() or 1
    val, step_s, nfr = cpu_time(cores, SAMPLE_BAND, SAMPLE_GOP, steps=1)
    print(json.dumps({"value": val, "unit": "frames/s", "cores": cores, "kind": "port",
                      "sample": sample_note(cores, nfr, SAMPLE_BAND, SAMPLE_GOP) + f"; {step_s:.1f} s; extrapolated, not a "
                      "full-config CPU run (see the reference arm's full_config_step)"}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu-sample"])
    ap.add_argument("--lanes", type=int, default=20, help="GOPs encoded in lock-step per GPU")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-decode", action="store_true")
    ap.add_argument("--skip-4k", action="store_true", help="N = 8: leave out the BASELINE configs[4] block")
    ap.add_argument("--skip-full", action="store_true", help="reference arm: leave out the full-config step")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "cpu-sample":
        run_cpu_sample()
        return

    import torch
    import torch.distributed as dist

    import basic_video_codec_b200 as bvc
    from basic_video_codec_b200 import sharding
    from basic_video_codec_b200._lib import measure_peaks

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    warmup = max(3, args.warmup)
    hbm_gbs, hbm_src = hbm_peak()
    frames, _pin = make_clip()
    ctx = bvc.Context(W, H, BS, R, QP, 1, False, False, IP, device=local_rank, max_lanes=args.lanes)
    out_buf_t = torch.empty(NFRAMES * W * H // 2, dtype=torch.uint8, pin_memory=True)
    out_buf = out_buf_t.numpy()

    # ---- issue-rate ceilings, measured on this GPU right before the timed region -------------------
    pk = measure_peaks(local_rank)

    # ---- value: clip resident in HBM (weak scaling: every rank its own 600 frames) -------------------
    ctx.clip_upload(frames)
    for _ in range(warmup):
        _, nbytes = ctx.encode_clip_resident(NFRAMES, out_buf)
    sha_resident = hashlib.sha256(out_buf[:nbytes].tobytes()).hexdigest()
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count()
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        _, nbytes = ctx.encode_clip_resident(NFRAMES, out_buf)
        dev_ms += ctx.last_kernel_times()[1]
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    dt = max_over_ranks(dt)
    dev_ms = max_over_ranks(dev_ms)
    value = world * NFRAMES * args.steps / dt
    if world > 1:   # every rank samples its own GPU; report the slowest clock and the union of throttle reasons
        ok = [c for c in gather(clocks) if c and c.get("sm_mhz")]
        if ok:
            clocks = {"sm_mhz": min(c["sm_mhz"] for c in ok), "sm_max_mhz": max(c["sm_max_mhz"] for c in ok),
                      "reasons": sorted({r for c in ok for r in c["reasons"]}), "samples": sum(c.get("samples", 0) for c in ok),
                      "per_rank_sm_mhz": [c["sm_mhz"] for c in ok]}

    # ---- e2e: host buffers through the public API ---------------------------------------------------
    e2e_bytes = ctx.encode_clip_into(frames, out_buf)  # warm-up of the host path
    sha_host = hashlib.sha256(out_buf[:e2e_bytes].tobytes()).hexdigest()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_bytes = ctx.encode_clip_into(frames, out_buf)
    barrier()
    dt_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_val = world * NFRAMES * e2e_steps / dt_e2e
    # the 600-frame stream: resident path == host path, and the same on every rank
    shas = gather((sha_resident, sha_host))
    stream_checks = {"resident_equals_host_path": all(a == b for a, b in shas),
                     "all_ranks_equal": len({a for a, _ in shas}) == 1, "sha256": sha_resident}
    if not (stream_checks["resident_equals_host_path"] and stream_checks["all_ranks_equal"]):
        raise SystemExit(f"bench: streams differ between paths / ranks: {shas}")
    # the CPU oracle's stream of this very clip, recorded once (oracle/gen_bench_clip_sha.py, 258 s on 8 cores): the number
    # below is a number for a bit-exact stream
    try:
        rec = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_clip_oracle.json")))
        stream_checks["equals_cpu_oracle_stream"] = (rec["sha256"] == sha_resident and rec["bytes"] == int(nbytes))
        stream_checks["cpu_oracle_sha256"] = rec["sha256"]
    except Exception as e:
        stream_checks["equals_cpu_oracle_stream"] = None
        stream_checks["cpu_oracle_sha256"] = f"record missing: {e}"
    if stream_checks["equals_cpu_oracle_stream"] is False:
        raise SystemExit("bench: the GPU stream differs from the CPU oracle's stream of the same clip")

    # ---- decoder, reported beside the headline: the stream just written, host buffers in and out ----
    decoder = None
    if not args.skip_decode:
        container_t = torch.empty(int(e2e_bytes), dtype=torch.uint8, pin_memory=True)   # pinned, like the encoder's input
        container = container_t.numpy()
        container[:] = out_buf[:e2e_bytes]
        dec_t = torch.empty((NFRAMES, H, W), dtype=torch.uint8, pin_memory=True)
        dec = dec_t.numpy()
        ctx.decode_clip(container, NFRAMES, out=dec)      # warm-up (allocations)
        barrier()
        t0 = time.perf_counter()
        DSTEPS = 2
        for _ in range(DSTEPS):
            ctx.decode_clip(container, NFRAMES, out=dec)
        barrier()
        dt_dec = max_over_ranks(time.perf_counter() - t0)
        # the link rate, measured in this run: the same 1.25 GB of pinned host memory filled by one device-to-host copy
        dsrc = torch.empty(dec_t.numel(), dtype=torch.uint8, device="cuda")
        dflat = dec_t.view(-1)
        dflat.copy_(dsrc, non_blocking=True)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        dflat.copy_(dsrc, non_blocking=True)
        ev1.record()
        torch.cuda.synchronize()
        link_s = ev0.elapsed_time(ev1) * 1e-3
        link_gbs = dec.nbytes / link_s / 1e9
        del dsrc
        decoder = {"value": world * NFRAMES * DSTEPS / dt_dec, "unit": "decoded frames/s", "ms_per_clip": dt_dec / DSTEPS * 1e3,
                   "h2d_bytes_per_step": int(e2e_bytes), "d2h_bytes_per_step": int(dec.nbytes),
                   "link_bound": {"ms": link_s * 1e3, "frac": link_s / (dt_dec / DSTEPS),
                                  "link_gbs": link_gbs,
                                  "note": "1.25 GB of decoded planes device -> pinned host as ONE copy, timed in this run; the decoder returns them "
                                          "step by step (30 strided copies: 52 GB/s in profiles/exp_pcie.py)"}}
        del dec_t, container_t

    # ---- roofline of the dominant kernel (motion estimation) ----------------------------------------
    # The timed steps above run two lane groups on separate streams (kernels of different groups overlap, so
    # event spans around a kernel are not exclusive).  Per-kernel launch durations are therefore taken from extra
    # passes over the same resident clip with one lane group: every kernel back to back on one stream, CUDA events
    # on that stream around each launch.
    groups = ctx.lane_groups
    ctx.clip_upload(frames)
    ctx.set_lane_groups(1)
    ctx.encode_clip_resident(NFRAMES, out_buf)
    kt_acc = None
    serial_ms = 0.0
    KSTEPS = 2
    for _ in range(KSTEPS):
        ctx.encode_clip_resident(NFRAMES, out_buf)
        kt, clip_ms = ctx.last_kernel_times()
        serial_ms += clip_ms / KSTEPS
        if kt_acc is None:
            kt_acc = {k: [0.0, 0] for k in kt}
        for k, (ms, n) in kt.items():
            kt_acc[k][0] += ms
            kt_acc[k][1] += n
    ctx.set_lane_groups(groups)
    me_ms, me_n = kt_acc["me"]
    me_ms_all = gather(me_ms / max(1, me_n))
    if world > 1:
        me_ms = max(me_ms_all) * me_n          # slowest rank
    px_per_launch = ctx.me_work_per_frame(1) * args.lanes        # one launch = frame k of every GOP lane
    me_avg_s = me_ms / max(1, me_n) * 1e-3
    achieved = px_per_launch / me_avg_s if me_avg_s > 0 else 0.0
    tq_ms, tq_n = kt_acc["tq_p"]
    tq_bytes = 5.0 * W * H * args.lanes                           # cur + pred in, int16 levels (coded, not stored) + recon out
    tq_gbs = tq_bytes / (tq_ms / max(1, tq_n) * 1e-3) / 1e9 if tq_ms > 0 else 0.0
    tq_dfma = tq_gbs * 1e9 / 5.0 * 32.0    # 5 algorithmic bytes and 32 DFMA per pixel
    share = {k: v[0] for k, v in kt_acc.items()}
    tot = sum(share.values()) or 1.0
    pk_all = gather(pk["px_absdiff_per_s"])
    peak_px = min(pk_all)                  # the slowest GPU's ceiling beside the slowest GPU's launch time
    sm_max = (clocks.get("sm_max_mhz") or 1965.0) * 1e6
    theo_px = 64.0 * 4.0 * pk["sm_count"] * sm_max     # 16 lanes/clk/SMSP x 4 SMSPs x 4 bytes per lane-op
    whole_step_px = ctx.me_work_per_frame(1) * (NFRAMES - NFRAMES // IP)

    # ---- strong scaling: the literal configs[3], one clip sharded over the ranks -----------------------
    ctx.close()
    ec = bvc.EncoderConfig(BS, R, IP, QP, nRefFrames=1)
    ngop = NFRAMES // IP
    strong = None
    with sharding.ShardedEncoder(ec, W, H, NFRAMES, rank=rank, world=world, device=local_rank,
                                 capacity=NFRAMES * W * H // 2) as enc:
        enc.upload(frames)
        for _ in range(warmup):
            res = enc.encode(resident=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = enc.encode(resident=True)
        barrier()
        dt_s = max_over_ranks(time.perf_counter() - t0)
        res = enc.encode(frames)             # warm-up of the host path
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res = enc.encode(frames)
        barrier()
        dt_se = max_over_ranks(time.perf_counter() - t0)
        if rank == 0:
            sha_strong = hashlib.sha256(res.tobytes()).hexdigest()
            ceiling = sharding.scaling_ceiling(ngop, world)
            strong = {"frames_per_s": NFRAMES * args.steps / dt_s, "ms_per_step": dt_s / args.steps * 1e3,
                      "e2e_frames_per_s": NFRAMES * e2e_steps / dt_se, "e2e_ms_per_step": dt_se / e2e_steps * 1e3,
                      "gops": ngop, "gops_on_busiest_rank": -(-ngop // world), "ceiling_speedup": ceiling,
                      "stream_bytes": int(res.size), "sha256": sha_strong,
                      "sha256_equals_single_gpu_stream": sha_strong == sha_resident,
                      "h2d_bytes_per_step_busiest_rank": int(-(-ngop // world) * IP * W * H),
                      "how": "basic_video_codec_b200.sharding.ShardedEncoder: contiguous GOP runs per rank, container fragment "
                             "lengths exchanged (one int64 per rank), fragments written by every rank into one shared pinned "
                             "host buffer, rank 0 returns the serial stream; frac_of_ceiling = frames_per_s / (single-GPU "
                             "frames_per_s x ceiling_speedup) is the driver's to compute from the N=1 line"}
            if not strong["sha256_equals_single_gpu_stream"]:
                raise SystemExit("bench: the sharded stream differs from the single-GPU stream")
        res = None

    line = {
        "metric": "encoded frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 SAD / int16 residual / f64 DCT", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_frames_per_step": world * NFRAMES,
                   "parallelism": f"GOP-sharded: {world} rank(s) x {args.lanes} GOP lanes in {groups} lane groups, no collective on the data path",
                   "l2": "inputs larger than L2 (1.25 GB clip per rank, 42 MB of planes per launch)"},
        "device_ms_per_step": dev_ms / args.steps,
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(frames.nbytes),
                "d2h_bytes_per_step": int(e2e_bytes + 16), "steps": e2e_steps},
        "strong": strong,
        "stream_checks": stream_checks,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "kernel": "me_tiled_kernel<16,4,8> (full-search SAD, VABSDIFF4.U8.ACC; 4 x 8 blocks per CTA, 4 x 4 in launches of fewer than four waves)", "bound": "int-simd",
            "achieved": achieved / 1e12, "peak": peak_px / 1e12, "unit": "Tpx-absdiff/s",
            "frac": achieved / peak_px,
            "peak_source": "measured in this run, right before the timed region: bvc_measure_peaks (dependent-free VABSDIFF4.U8.ACC "
                           "chains on every SM, best of 5 launches, CUDA events) x 4 bytes per lane-op",
            "peak_theoretical": theo_px / 1e12, "frac_of_theoretical": achieved / theo_px,
            "peak_theoretical_source": "64 lanes/clk/SM x SMs x max SM clock x 4",
            "per_rank_peak": [p / 1e12 for p in pk_all] if world > 1 else None,
            "whole_step_frac": (whole_step_px / (dt / args.steps)) / peak_px,
            "traffic": ME_TRAFFIC_BYTES_PER_LANE * args.lanes,
            "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one 10-lane launch / 10 (profiles/r2_ncu_me_kernel.csv); "
                              "algorithmic: 2 planes of 2.09 MB per lane",
            "algorithmic_bytes": 2 * W * H * args.lanes + 16 * (W // BS) * (H // BS) * args.lanes,
            "launch_ms": me_avg_s * 1e3, "launches": me_n,
            "per_rank_launch_ms": me_ms_all if world > 1 else None,
            "gpos_per_s": achieved / (BS * BS) / 1e9, "share_of_kernel_time": share["me"] / tot,
        },
        "roofline_transform": {
            "kernel": "tq_pframe_kernel<16> (residual+DCT+quant+IDCT+recon+entropy, fp64)", "bound": "hbm",
            "achieved": tq_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": tq_gbs / hbm_gbs,
            "peak_source": hbm_src, "traffic": TQ_TRAFFIC_BYTES_PER_LANE * args.lanes,
            "share_of_kernel_time": share["tq_p"] / tot,
            "note": "fp64-pipe / issue bound (32 DFMA/px), not HBM bound: see DESIGN.md",
            # the same launches against the fp64 pipe: 32 DFMA per pixel (folded separable forward + inverse transform)
            "fp64": {"achieved": tq_dfma, "peak": pk["dfma_thread_ops_per_s"], "unit": "DFMA thread-ops/s",
                     "frac": tq_dfma / pk["dfma_thread_ops_per_s"], "algorithmic": "32 DFMA per pixel",
                     "peak_source": "measured in this run (bvc_measure_peaks, dependent-free fma.rn.f64 chains)"},
        },
        "decoder": decoder,
        # latency-bound kernel (no roofline fraction, time per diagonal of the anti-diagonal wavefront)
        "intra_wavefront": {"ms_per_I_step": kt_acc["tq_i"][0] / KSTEPS, "diagonals": (W // BS) + (H // BS) - 1,
                            "us_per_diagonal": kt_acc["tq_i"][0] / KSTEPS * 1e3 / ((W // BS) + (H // BS) - 1),
                            "note": f"{args.lanes} I frames in one launch (one warp per block pair at this lane count) + their entropy-coding kernel"},
        "kernel_ms_per_step": {k: v[0] / KSTEPS for k, v in kt_acc.items()},
        "kernel_timing": f"{KSTEPS} extra passes with one lane group (kernels serialised on one stream, {serial_ms:.2f} ms per pass); "
                         "the timed steps overlap kernel tails across lane groups",
        "bitstream_bytes_per_step": int(nbytes),
    }

    # ---- BASELINE configs[4] (4K, 64 GOPs, r=64, 4 references, "whole 8 x B200 box"): measured when the whole box is there ----
    if world == 8 and not args.skip_4k:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("run_c4_8gpu", os.path.join(ROOT, "profiles", "run_c4_8gpu.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            c5 = mod.measure(rank, world, local_rank)
        except Exception as e:   # a secondary figure, not a reason to lose the line -- but every rank must take the same path
            c5 = {"error": repr(e)}
        line["config5_4k_whole_box"] = c5

    if rank == 0 and world == 1 and not args.skip_cpu:
        # in a subprocess: this process maps libbvc_b200.so only
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "cpu-sample"], capture_output=True, text=True, timeout=300)
            line["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:   # the baseline is a reported figure, not a reason to lose the line
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
