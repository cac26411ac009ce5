#!/usr/bin/env python3
"""Generate the rate-control lookup tables shipped in basic_video_codec_b200/encoder/RateControl/lookups/.

Same file naming and CSV layout as the reference's tables (encoder/RateControl/lookups/<W>_<H>_<i>_<I|P>.csv:
a header row of QPs 0..11, one row of average bits per block row), measured here with the CPU oracle on the
synthetic CIF / QCIF stand-in clip (Foreman, which the reference's tables were measured on, is a git-LFS pointer
and is not available).  Run:  python -m oracle.gen_rc_lookups
TEST / TOOLING INFRASTRUCTURE: the product only reads the CSV files."""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import bindings as ob  # noqa: E402
from tests import synth  # noqa: E402
from tests.golden_util import split_container  # noqa: E402

OUT = os.path.join(ROOT, "basic_video_codec_b200", "encoder", "RateControl", "lookups")


def main():
    os.makedirs(OUT, exist_ok=True)
    for (W, H) in ((352, 288), (176, 144)):
        frames = synth.moving_clip(352 + W, H, W, 8, step=3, clamp=24)
        for bs in (8, 16):
            rows = H // bs
            lg = bs.bit_length() - 1
            for kind, ip in (("I", 1), ("P", 8)):
                qps, vals = [], []
                for qp in range(0, 12):
                    if qp > lg + 7:
                        continue
                    cfg = ob.make_config(W, H, bs, 4, qp, nref=1, i_period=ip)
                    data, _ = ob.encode_clip(cfg, frames, nthreads=8, want_recon=False)
                    sel = [6 + len(p) + len(c) for m, p, c in split_container(data) if (m == 1) == (kind == "I")]
                    qps.append(qp)
                    vals.append(round(sum(sel) * 8 / (len(sel) * rows)))
                path = os.path.join(OUT, f"{W}_{H}_{bs}_{kind}.csv")
                with open(path, "w", newline="") as fh:
                    w = csv.writer(fh)
                    w.writerow(qps)
                    w.writerow(vals)
                print(path, vals)


if __name__ == "__main__":
    main()
