#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED Python reference at /root/reference.

Run in the build container only:  python -m oracle.gen_golden
Each fixture stores the input planes, the parameters, and the outputs of the reference's own
encode_video() run with the defined fp64 DCT monkeypatched in (oracle/ref_harness.py):
container bytes, reconstructed planes, quantised level planes, mv.txt, plus per-frame details from
the in-memory frame loop (MVs, modes, avg_mae, comparison counts, bits per row, bit strings).
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from tests import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (generator, kwargs for the generator, encoder parameters)
    "fs_i8_r4_qp3": ("moving", dict(seed=11, height=64, width=96, nframes=10, step=3, clamp=16),
                     dict(block=8, search_range=4, qp=3, i_period=8, nref=1)),
    "fs_i16_r2_nref4": ("moving", dict(seed=12, height=64, width=96, nframes=7, step=2, clamp=16),
                        dict(block=16, search_range=2, qp=2, i_period=8, nref=4)),
    "fastme_i16_nref4": ("moving", dict(seed=13, height=64, width=96, nframes=8, step=2, clamp=24),
                         dict(block=16, search_range=4, qp=3, i_period=8, nref=4, fastme=True)),
    "frac_fs_i8_r2_nref2": ("moving", dict(seed=14, height=48, width=64, nframes=6, step=2, clamp=16),
                            dict(block=8, search_range=2, qp=1, i_period=4, nref=2, frac=True)),
    "frac_fastme_i8_nref3": ("moving", dict(seed=15, height=48, width=64, nframes=7, step=2, clamp=16),
                             dict(block=8, search_range=2, qp=4, i_period=5, nref=3, frac=True, fastme=True)),
    "fs_i4_r3_nref2": ("moving", dict(seed=16, height=32, width=48, nframes=5, step=2, clamp=16),
                       dict(block=4, search_range=3, qp=1, i_period=3, nref=2)),
    "ties_i8_r4": ("poster", dict(seed=17, height=64, width=96, nframes=5),
                   dict(block=8, search_range=4, qp=2, i_period=5, nref=2)),
    "ties_fastme_i16": ("poster", dict(seed=18, height=64, width=96, nframes=5),
                        dict(block=16, search_range=4, qp=5, i_period=5, nref=2, fastme=True)),
    "fs_i16_r8_qp6": ("moving", dict(seed=19, height=96, width=128, nframes=4, step=6, clamp=32),
                      dict(block=16, search_range=8, qp=6, i_period=4, nref=1)),
}


def frame_details(frames, enc):
    objs = rh.ref_encode_frames(frames, **enc)
    det = []
    for fr in objs:
        d = {"intra": int(fr.prediction_mode.value), "avg_mae": float(fr.avg_mae),
             "mae_comparisons": int(fr.total_mae_comparisons),
             "bits_per_row": [int(b) for b in fr.bits_per_row],
             "pred_nbits": len(fr.entropy_encoded_prediction_data),
             "coef_nbits": len(fr.entropy_encoded_DCT_coffs),
             "pred_sha": hashlib.sha256(fr.entropy_encoded_prediction_data.tobytes()).hexdigest(),
             "coef_sha": hashlib.sha256(fr.entropy_encoded_DCT_coffs.tobytes()).hexdigest()}
        if d["intra"]:
            d["modes"] = [int(m) for m in fr.intra_modes]
        else:
            keys = sorted(fr.mv_field.keys(), key=lambda k: (k[1], k[0]))
            d["mv"] = [[int(v) for v in fr.mv_field[k]] for k in keys]
        det.append(d)
    return det


RC_CASES = {
    # name: (gen kwargs, enc kwargs, RCflag, targetBR)
    "rc1_i16": (dict(seed=51, height=64, width=96, nframes=7, step=3, clamp=16), dict(block=16, search_range=4, qp=4, i_period=4, nref=1), 1, 280000),
    "rc1_i8_nref2": (dict(seed=52, height=48, width=64, nframes=6, step=2, clamp=16), dict(block=8, search_range=2, qp=3, i_period=3, nref=2), 1, 220000),
    "rc2_i16": (dict(seed=53, height=64, width=96, nframes=7, step=3, clamp=16), dict(block=16, search_range=4, qp=5, i_period=7, nref=1), 2, 200000),
    "rc3_i8_scene": (dict(seed=54, height=48, width=64, nframes=6, step=1, clamp=16, noise=0), dict(block=8, search_range=2, qp=5, i_period=6, nref=2), 3, 330000),
    "rc1_frac_fastme": (dict(seed=55, height=48, width=64, nframes=5, step=2, clamp=16), dict(block=8, search_range=2, qp=3, i_period=5, nref=2, frac=True, fastme=True), 1, 150000),
}


def rc_main(only=None):
    from oracle import rc_oracle
    ns = rh.load_reference()
    for name, (gk, enc, rcflag, br) in RC_CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        frames = synth.moving_clip(**gk)
        if name == "rc3_i8_scene":   # a hard cut in the middle of the GOP so the scene-change path triggers
            frames[3:] = synth.moving_clip(seed=99, height=gk["height"], width=gk["width"], nframes=gk["nframes"] - 3, step=1, clamp=16, noise=0, blur=3)
        table = rc_oracle.measure_table(frames, enc["block"], nref=1)
        out = ref_encode_video_rc(ns, frames, enc, rcflag, br, table)
        meta = {"generator": "moving", "gen_kwargs": gk, "enc": enc, "rcflag": rcflag, "targetBR": br,
                "table": {str(k): v for k, v in table.items()}, "dct_mode": "fp64_defined",
                "encoded_sha256": hashlib.sha256(out["encoded"]).hexdigest()}
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), frames=frames,
                            encoded=np.frombuffer(out["encoded"], dtype=np.uint8), recon=out["recon"],
                            meta=np.array(json.dumps(meta)))
        print(f"{name}: {len(out['encoded'])} B  ({time.time() - t0:.1f}s)")


def _y_generator_frames(width, height, n):
    """The reference's own synthetic source (tests/y_generator.py), the stand-in for Foreman (an LFS pointer here)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_y_generator", os.path.join(rh.REFERENCE_ROOT, "tests", "y_generator.py"))
    yg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(yg)
    raw = yg.generate_yuv_bytestream(width, height, n)
    return np.frombuffer(raw, dtype=np.uint8).reshape(n, height, width).copy()


def cif_main(only=None):
    """BASELINE configs 2 and 3 at their real geometry (CIF), input from the reference's tests/y_generator.py:
      cif_c2      i=16, FastME, nRefFrames=4, 24 frames, I_Period 8                 (assign2/FastME.py:8-11 path)
      rc1_cif_c3  i=16, r=4, half-pel, RCflag=1, 21 frames, I_Period 21, 2.4 Mbit/s (assign3/Ex1.py:16-31 path), with the
                  reference's OWN lookup tables (encoder/RateControl/lookups/352_288_16_{I,P}.csv through its own loader)
    Outputs of the reference's encode_video (defined fp64 DCT): container bytes + reconstruction hash + per-frame sizes."""
    ns = rh.load_reference()
    cases = {
        "cif_c2": (24, dict(block=16, search_range=16, qp=3, i_period=8, nref=4, fastme=True), 0, 0),
        "rc1_cif_c3": (21, dict(block=16, search_range=4, qp=4, i_period=21, nref=1, frac=True), 1, 2_400_000),
    }
    for name, (n, enc, rcflag, br) in cases.items():
        if only and name not in only:
            continue
        t0 = time.time()
        frames = _y_generator_frames(352, 288, n)
        table = None
        if rcflag:
            ec0 = rh.make_config(ns, block=enc["block"], search_range=enc["search_range"], qp=enc["qp"], i_period=enc["i_period"],
                                 width=352, height=288)
            table = ns.encoder.get_combined_lookup_table(ns.encoder.rc_lookup_file_path(ec0, "I"), ns.encoder.rc_lookup_file_path(ec0, "P"))
            out = ref_encode_video_rc(ns, frames, enc, rcflag, br, None)
        else:
            out = rh.ref_encode_video(frames, **enc)
        sizes, o, data = [], 0, out["encoded"]
        while o < len(data):
            pl = int.from_bytes(data[o + 1:o + 3], "big")
            cl = int.from_bytes(data[o + 3 + pl:o + 6 + pl], "big")
            sizes.append([int(data[o]), pl, cl])
            o += 6 + pl + cl
        meta = {"generator": f"reference tests/y_generator.py generate_yuv_bytestream(352,288,{n})", "enc": enc, "rcflag": rcflag,
                "targetBR": br, "dct_mode": "fp64_defined",
                "table": ({str(k): v for k, v in table.items()} if table else None),
                "table_source": "reference encoder/RateControl/lookups/352_288_16_{I,P}.csv via its own get_combined_lookup_table" if table else None,
                "encoded_sha256": hashlib.sha256(out["encoded"]).hexdigest(),
                "recon_sha256": hashlib.sha256(out["recon"].tobytes()).hexdigest(), "frame_records": sizes}
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), frames=frames,
                            encoded=np.frombuffer(out["encoded"], dtype=np.uint8), meta=np.array(json.dumps(meta)))
        print(f"{name}: {len(out['encoded'])} B  ({time.time() - t0:.1f}s)")


def ref_encode_video_rc(ns, frames, enc, rcflag, br, table):
    """The reference's own encode_video with RCflag set and `table` returned by its lookup loader (table=None: the
    reference loads its own CSVs)."""
    import tempfile
    rh.set_dct_mode("fp64_defined")
    n, H, W = frames.shape
    ec = ns.params.EncoderConfig(enc["block"], enc["search_range"], enc["i_period"], enc["qp"], nRefFrames=enc.get("nref", 1),
                                 fastME=enc.get("fastme", False), fracMeEnabled=enc.get("frac", False), RCflag=rcflag,
                                 targetBR=br, resolution=(W, H))
    orig = ns.encoder.get_combined_lookup_table
    if table is not None:
        ns.encoder.get_combined_lookup_table = lambda a, b: {int(k): dict(v) for k, v in table.items()}
    try:
        with tempfile.TemporaryDirectory(prefix="bvc_ref_rc_") as td:
            yfile = os.path.join(td, "clip.y")
            open(yfile, "wb").write(np.ascontiguousarray(frames).tobytes())
            params = ns.input_parameters.InputParameters(yfile, W, H, ec, frames_to_process=n)
            ns.encoder.encode_video(params)
            fio = ns.encoder.FileIOHelper(params)
            return {"encoded": open(fio.get_encoded_file_name(), "rb").read(),
                    "recon": np.fromfile(fio.get_mc_reconstructed_file_name(), dtype=np.uint8).reshape(n, H, W)}
    finally:
        ns.encoder.get_combined_lookup_table = orig


def main(only=None):
    os.makedirs(GOLD, exist_ok=True)
    for name, (gen, gk, enc) in CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        frames = synth.moving_clip(**gk) if gen == "moving" else synth.posterised_clip(**gk)
        out = rh.ref_encode_video(frames, **enc)
        det = frame_details(frames, enc)
        meta = {"generator": gen, "gen_kwargs": gk, "enc": enc, "frames": det,
                "dct_mode": "fp64_defined", "encoded_sha256": hashlib.sha256(out["encoded"]).hexdigest()}
        np.savez_compressed(os.path.join(GOLD, name + ".npz"),
                            frames=frames, encoded=np.frombuffer(out["encoded"], dtype=np.uint8),
                            recon=out["recon"], levels=out["levels"], resid_mc=out["resid_mc"],
                            resid_nomc=out["resid_nomc"], mv_txt=np.array(out["mv_txt"]),
                            meta=np.array(json.dumps(meta)))
        print(f"{name}: {len(out['encoded'])} B  ({time.time() - t0:.1f}s)")

    # ---- rate control (RCflag 1/2/3) with a lookup table measured by the oracle, patched into the reference ----
    rc_main(only)
    # ---- BASELINE configs 2 and 3 at CIF (reference generator input, reference lookup tables) ----
    cif_main(only)

    # CIF stand-in for BASELINE config 1 (Foreman is an LFS pointer): the reference's own synthetic
    # generator tests/y_generator.py, 10 frames, i=8 r=4 qp=3 I_Period=8.
    if not only or "cif_c1" in only:
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_y_generator",
                                                      os.path.join(rh.REFERENCE_ROOT, "tests", "y_generator.py"))
        yg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(yg)
        raw = yg.generate_yuv_bytestream(352, 288, 10)
        frames = np.frombuffer(raw, dtype=np.uint8).reshape(10, 288, 352)
        enc = dict(block=8, search_range=4, qp=3, i_period=8, nref=1)
        t0 = time.time()
        out = rh.ref_encode_video(frames, **enc)
        meta = {"generator": "reference tests/y_generator.py generate_yuv_bytestream(352,288,10)", "enc": enc,
                "dct_mode": "fp64_defined", "encoded_sha256": hashlib.sha256(out["encoded"]).hexdigest(),
                "recon_sha256": hashlib.sha256(out["recon"].tobytes()).hexdigest(),
                "levels_sha256": hashlib.sha256(out["levels"].tobytes()).hexdigest()}
        np.savez_compressed(os.path.join(GOLD, "cif_c1.npz"), frames=frames,
                            encoded=np.frombuffer(out["encoded"], dtype=np.uint8),
                            mv_txt=np.array(out["mv_txt"]), meta=np.array(json.dumps(meta)))
        print(f"cif_c1: {len(out['encoded'])} B  ({time.time() - t0:.1f}s)")


if __name__ == "__main__":
    main(sys.argv[1:])
