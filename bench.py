#!/usr/bin/env python3
"""bench.py -- headline benchmark: encoded frames/s on synthetic 1920x1088 luma, 600 frames, i=16, r=32
integer full search, I_Period=30 (BASELINE.json configs[3]), GOP-sharded: one process per GPU, every
rank encodes its own 20 GOPs (weak scaling, no data-path collective; NCCL only for barrier / max).

  python bench.py [--gpus N] [--steps K] [--warmup W]              our CUDA path (libbvc_b200.so)
  python bench.py --impl reference [...]                           CPU arm: the oracle port on all host cores

One "step" = one pass of the encoder hot path over the rank's whole 600-frame clip.
  value : frames/s with the clip already resident in HBM (bitstreams still come back to the host)
  e2e   : frames/s through the public API with the clip in pinned HOST memory (H2D inside the timed region)
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, BS, R, QP, IP, NFRAMES = 1920, 1088, 16, 32, 4, 30, 600
# DRAM traffic per lane (one 1080p frame) of a launch, from the committed ncu --set full captures (profiles/r1_ncu_*.csv)
ME_TRAFFIC_BYTES_PER_LANE = (41.834240e6 + 0.472576e6) / 10       # 10-lane launch (two lane groups)
TQ_TRAFFIC_BYTES_PER_LANE = (43.469568e6 + 3.402240e6) / 10
WORKLOAD = "synthetic 1920x1088 Y plane, 600 frames, i=16, r=32 full-search, I_Period=30, nRefFrames=1, QP=4 (BASELINE configs[3])"


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            peaks["hbm_gbs"] = float(json.load(open(p))["hbm_gbs"])
            peaks["hbm_src"] = "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    q = os.path.join(ROOT, "profiles", "int_simd_peak.json")
    d = json.load(open(q))
    peaks["px_per_s"] = float(d["px_absdiff_per_s"])
    peaks["int_src"] = d["how"]
    f = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
    peaks["dfma_per_s"] = float(f["dfma_thread_ops_per_s"])
    peaks["fp64_src"] = f["how"]
    return peaks


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


def make_clip(rank, pinned=True):
    from tests import synth
    if pinned:
        import torch
        buf = torch.empty((NFRAMES, H, W), dtype=torch.uint8, pin_memory=True)
        arr = buf.numpy()
    else:
        buf = None
        arr = np.empty((NFRAMES, H, W), np.uint8)
    # one texture + random walk per rank; generated GOP by GOP to bound peak memory
    arr[:] = synth.moving_clip(1080 + rank, H, W, NFRAMES, step=6, clamp=96, noise=2)
    return arr, buf


def cpu_sample(rank, cores, budget_s):
    """Bounded CPU sample of the same workload: `cores` GOPs truncated to I + 5 P, one GOP per thread (the workload's GOPs
    are I + 29 P and a P frame costs ~15x an I frame, so the sample's frames are ~13 % cheaper on average than the
    workload's: the CPU figure is slightly optimistic).  If that exceeds the budget, the planes are cropped to a band of
    block rows (full width) and the result is scaled by band_height / 1088."""
    from oracle import bindings as ob
    from tests import synth
    GOP_SAMPLE = 6
    per_thread_full = 0.19 + (GOP_SAMPLE - 1) * 3.0   # s, measured on the build container (I + 5 P at 1080p r=32)
    band_h = H
    if per_thread_full > budget_s:
        rows = max(6, int((H // BS) * budget_s / per_thread_full))
        band_h = rows * BS
    clip = synth.moving_clip(4242 + rank, H, W, GOP_SAMPLE, step=6, clamp=96, noise=2)
    gops = []
    for g in range(cores):
        c = np.roll(clip, shift=7 * g, axis=2)[:, :band_h, :]
        gops.append(c)
    frames = np.ascontiguousarray(np.concatenate(gops, axis=0))
    cfg = ob.make_config(W, band_h, BS, R, QP, nref=1, i_period=GOP_SAMPLE)
    return ob, cfg, frames, band_h


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    budget = max(1.5, min(8.0, 150.0 / max(1, args.steps + args.warmup)))
    ob, cfg, frames, band_h = cpu_sample(rank, cores, budget)
    nfr = frames.shape[0]
    for _ in range(max(0, args.warmup)):
        ob.encode_clip(cfg, frames, nthreads=cores, want_recon=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ob.encode_clip(cfg, frames, nthreads=cores, want_recon=False)
    dt = time.perf_counter() - t0
    eq_frames = nfr * band_h / H
    val = eq_frames * args.steps / dt
    sample = (f"{cores} GOPs x (I + 5 P) of the workload, one GOP per thread, band of {band_h}/{H} luma rows at full width; "
              f"frames/s = {nfr} frames x {band_h}/{H} per step")
    line = {
        "impl": "reference", "metric": "encoded frames/s", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8 SAD / int16 residual / f64 DCT", "data": "synthetic",
        "config": {"workload": WORKLOAD, "parallelism": f"{cores} host threads (OpenMP over GOPs)"},
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of the reference's algorithm (oracle/bvc_oracle.c, -O3 AVX2). The reference itself is pure "
                "Python/NumPy and ~1000x slower: ~440 s per 1080p r=32 P frame on one core (BASELINE.md)",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lanes", type=int, default=20, help="GOPs encoded in lock-step per GPU")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-decode", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import basic_video_codec_b200 as bvc

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

    peaks = load_peaks()
    frames, _pin = make_clip(rank)
    ctx = bvc.Context(W, H, BS, R, QP, 1, False, False, IP, device=local_rank, max_lanes=args.lanes)
    out_buf_t = torch.empty(NFRAMES * W * H // 2, dtype=torch.uint8, pin_memory=True)
    out_buf = out_buf_t.numpy()

    # ---- value: clip resident in HBM --------------------------------------------------------------
    ctx.clip_upload(frames)
    for _ in range(max(3, args.warmup)):
        _, nbytes = ctx.encode_clip_resident(NFRAMES, out_buf)
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
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        ok = [c for c in allc if c and c.get("sm_mhz")]
        if ok:
            clocks = {"sm_mhz": min(c["sm_mhz"] for c in ok), "sm_max_mhz": max(c["sm_max_mhz"] for c in ok),
                      "reasons": sorted({r for c in ok for r in c["reasons"]}), "samples": sum(c.get("samples", 0) for c in ok),
                      "per_rank_sm_mhz": [c["sm_mhz"] for c in ok]}

    # ---- e2e: host buffers through the public API ---------------------------------------------------
    ctx.encode_clip_into(frames, out_buf)  # warm-up of the host path
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_bytes = ctx.encode_clip_into(frames, out_buf)
    barrier()
    dt_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_val = world * NFRAMES * e2e_steps / dt_e2e

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
        decoder = {"value": world * NFRAMES * DSTEPS / dt_dec, "unit": "decoded frames/s", "ms_per_clip": dt_dec / DSTEPS * 1e3,
                   "h2d_bytes_per_step": int(e2e_bytes), "d2h_bytes_per_step": int(dec.nbytes),
                   "note": "bvc_decode_clip on the container written above, 1.25 GB of decoded planes returned over PCIe (23 ms at the link rate) after 5.6 ms of upload and tokenizing; see DESIGN.md D1-D6"}
        del dec_t, container_t

    # ---- roofline of the dominant kernel (motion estimation) ----------------------------------------
    # The timed steps above run two lane groups on separate streams (kernels of different groups overlap, so
    # event spans around a kernel are not exclusive).  Per-kernel launch durations are therefore taken from extra
    # passes over the same resident clip with one lane group: every kernel back to back on one stream, CUDA events
    # on that stream around each launch.
    groups = ctx.lane_groups
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
    me_ms_all = [None] * world
    if world > 1:
        dist.all_gather_object(me_ms_all, me_ms / max(1, me_n))
        me_ms, me_n = max(me_ms_all) * me_n, me_n          # slowest rank
    px_per_launch = ctx.me_work_per_frame(1) * args.lanes        # one launch = frame k of every GOP lane
    me_avg_s = me_ms / max(1, me_n) * 1e-3
    achieved = px_per_launch / me_avg_s if me_avg_s > 0 else 0.0
    tq_ms, tq_n = kt_acc["tq_p"]
    tq_bytes = 5.0 * W * H * args.lanes                           # cur + pred in, int16 levels (coded, not stored) + recon out
    tq_gbs = tq_bytes / (tq_ms / max(1, tq_n) * 1e-3) / 1e9 if tq_ms > 0 else 0.0
    tq_dfma = tq_gbs * 1e9 / 5.0 * 32.0    # 5 algorithmic bytes and 32 DFMA per pixel
    share = {k: v[0] for k, v in kt_acc.items()}
    tot = sum(share.values()) or 1.0

    line = {
        "metric": "encoded frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 SAD / int16 residual / f64 DCT", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_frames_per_step": world * NFRAMES,
                   "parallelism": f"GOP-sharded: {world} rank(s) x {args.lanes} GOP lanes in {groups} lane groups, no collective on the data path",
                   "l2": "inputs larger than L2 (1.25 GB clip per rank, 42 MB of planes per launch)"},
        "device_ms_per_step": dev_ms / args.steps,
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(frames.nbytes),
                "d2h_bytes_per_step": int(e2e_bytes + 16), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "kernel": "me_tiled_kernel<16,4> (full-search SAD, VABSDIFF4.U8.ACC)", "bound": "int-simd",
            "achieved": achieved / 1e12, "peak": peaks["px_per_s"] / 1e12, "unit": "Tpx-absdiff/s",
            "frac": achieved / peaks["px_per_s"], "traffic": ME_TRAFFIC_BYTES_PER_LANE * args.lanes,
            "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one 10-lane launch / 10 (profiles/r1_ncu_me_kernel.csv); "
                              "algorithmic: 2 planes of 2.09 MB per lane",
            "algorithmic_bytes": 2 * W * H * args.lanes + 16 * (W // BS) * (H // BS) * args.lanes,
            "peak_source": peaks["int_src"], "launch_ms": me_avg_s * 1e3, "launches": me_n,
            "per_rank_launch_ms": me_ms_all if world > 1 else None,
            "gpos_per_s": achieved / (BS * BS) / 1e9, "share_of_kernel_time": share["me"] / tot,
        },
        "roofline_transform": {
            "kernel": "tq_pframe_kernel<16> (residual+DCT+quant+IDCT+recon+entropy, fp64)", "bound": "hbm",
            "achieved": tq_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": tq_gbs / peaks["hbm_gbs"],
            "peak_source": peaks["hbm_src"], "traffic": TQ_TRAFFIC_BYTES_PER_LANE * args.lanes,
            "share_of_kernel_time": share["tq_p"] / tot,
            "note": "fp64-pipe bound (32 DFMA/px), not HBM bound: see DESIGN.md",
            # the same launches against the fp64 pipe: 32 DFMA per pixel (folded separable forward + inverse transform)
            "fp64": {"achieved": tq_dfma, "peak": peaks["dfma_per_s"], "unit": "DFMA thread-ops/s",
                     "frac": tq_dfma / peaks["dfma_per_s"], "algorithmic": "32 DFMA per pixel", "peak_source": peaks["fp64_src"]},
        },
        "decoder": decoder,
        "kernel_ms_per_step": {k: v[0] / KSTEPS for k, v in kt_acc.items()},
        "kernel_timing": f"{KSTEPS} extra passes with one lane group (kernels serialised on one stream, {serial_ms:.2f} ms per pass); "
                         "the timed steps overlap kernel tails across lane groups",
        "bitstream_bytes_per_step": int(nbytes),
    }

    if rank == 0 and world == 1 and not args.skip_cpu:
        cores = os.cpu_count() or 1
        ob, cfg, sframes, band_h = cpu_sample(rank, cores, 8.0)
        t0 = time.perf_counter()
        ob.encode_clip(cfg, sframes, nthreads=cores, want_recon=False)
        dtc = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": sframes.shape[0] * band_h / H / dtc, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{cores} GOPs x (I + 5 P) of the workload, one GOP per host thread, band of {band_h}/{H} luma rows "
                      f"({sframes.shape[0]} frames, {dtc:.1f} s); "
                      f"C restatement oracle/bvc_oracle.c; the Python reference itself needs ~440 s per P frame (BASELINE.md)",
        }
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
