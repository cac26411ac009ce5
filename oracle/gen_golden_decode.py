#!/usr/bin/env python3
"""Pin the decoder: run the UNMODIFIED reference decode_video() (decoder.py:26-87, imported from /root/reference
through oracle/ref_harness.py, fp64_defined DCT) on the container bytes of every committed golden and record
what it produced (sha256 of mc_decoded.yuv, frame count, whether it equals the encoder's reconstruction) in
tests/golden/decode_ref.json.  Build container only; the GPU box reads the JSON.

Usage: python -m oracle.gen_golden_decode [name ...]
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from tests import golden_util as gu  # noqa: E402

OUT = os.path.join(gu.GOLD, "decode_ref.json")


def ref_decode_video(ns, encoded: bytes, recon: np.ndarray, enc: dict, rcflag=0, target_br=0):
    """The reference's decode_video on `encoded`; returns the decoded planes."""
    rh.set_dct_mode("fp64_defined")
    n, H, W = recon.shape
    ec = ns.params.EncoderConfig(enc["block"], enc["search_range"], enc["i_period"], enc["qp"], nRefFrames=enc.get("nref", 1),
                                 fastME=enc.get("fastme", False), fracMeEnabled=enc.get("frac", False), RCflag=rcflag,
                                 targetBR=target_br, resolution=(W, H))
    with tempfile.TemporaryDirectory(prefix="bvc_ref_dec_") as td:
        yfile = os.path.join(td, "clip.y")
        open(yfile, "wb").close()
        params = ns.input_parameters.InputParameters(yfile, W, H, ec, frames_to_process=n)
        fio = ns.decoder.FileIOHelper(params)
        open(fio.get_encoded_file_name(), "wb").write(encoded)
        open(fio.get_mc_reconstructed_file_name(), "wb").write(recon.tobytes())   # only read for the PSNR log line
        ns.decoder.decode_video(params)
        return np.fromfile(fio.get_mc_decoded_file_name(), dtype=np.uint8).reshape(-1, H, W)


def main(only=None):
    ns = rh.load_reference()
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in gu.names():
        if only and name not in only:
            continue
        g = gu.load(name)
        meta = g["meta"]
        n, H, W = g["frames"].shape
        # goldens that store only the hash of the reconstruction: the reference decoder reads mc_reconstructed.yuv for a
        # PSNR log line only, so zeros do; the decoded planes are then compared with the stored hash
        recon = g["recon"] if "recon" in g else np.zeros((n, H, W), np.uint8)
        t0 = time.time()
        dec = ref_decode_video(ns, g["encoded"], recon, meta["enc"], meta.get("rcflag", 0), meta.get("targetBR", 0))
        res[name] = {"frames": int(dec.shape[0]), "decoded_sha256": hashlib.sha256(dec.tobytes()).hexdigest(),
                     "equals_encoder_recon": bool(np.array_equal(dec, recon)) if "recon" in g
                     else hashlib.sha256(dec.tobytes()).hexdigest() == meta["recon_sha256"]}
        print(f"{name}: {res[name]}  ({time.time() - t0:.1f}s)", flush=True)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:] or None)
