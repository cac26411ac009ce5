#!/usr/bin/env python3
"""Take over the reference's rate-control lookup tables (encoder/RateControl/lookups/<W>_<H>_<i>_<I|P>.csv: average bits
per block row for every QP, measured by the reference's authors on their CIF / QCIF sequences) as the tables this package
ships.  They are data a drop-in has to match: with RCflag != 0 the reference picks every row's QP from them
(encoder/RateControl/RateControl.py:34-43), so different tables mean different streams.  Numbers only, re-written
through the csv module; the tables measured on this repo's synthetic stand-in (oracle/gen_rc_lookups.py) remain
available for resolutions the reference has none for.  Build container only.
Usage: python -m oracle.import_rc_lookups"""
import csv
import glob
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(os.environ.get("BVC_REFERENCE_ROOT", "/root/reference"), "encoder", "RateControl", "lookups")
DST = os.path.join(ROOT, "basic_video_codec_b200", "encoder", "RateControl", "lookups")

if __name__ == "__main__":
    for path in sorted(glob.glob(os.path.join(SRC, "*.csv"))):
        rows = [r for r in csv.reader(open(path, newline="")) if r]
        qps = [int(float(x)) for x in rows[0]]
        vals = [int(float(x)) for x in rows[1]]
        with open(os.path.join(DST, os.path.basename(path)), "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(qps)
            w.writerow(vals)
        print(os.path.basename(path), dict(zip(qps, vals)))
