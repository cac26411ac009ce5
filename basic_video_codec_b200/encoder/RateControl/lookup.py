"""Row-size lookup tables (reference encoder/RateControl/lookup.py).

File naming and CSV layout are the reference's (`<W>_<H>_<i>_<I|P>.csv`: a header row of QPs, one row of
average bits per block row), so its tables can be dropped into the lookup directory unchanged.  The
directory is `$BVC_RC_LOOKUP_DIR` if set, else `lookups/` next to this file, which ships tables measured
with this encoder on the synthetic CIF/QCIF stand-in clip (Foreman is not redistributable here).
"""
import csv
import os
import re


def lookup_dir():
    return os.environ.get("BVC_RC_LOOKUP_DIR") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lookups")


def rc_lookup_file_path(ec, i_period_str=None):
    if not i_period_str:
        i_period_str = "I" if ec.I_Period == 1 else "P"
    return os.path.join(lookup_dir(), f"{ec.resolution[0]}_{ec.resolution[1]}_{ec.block_size}_{i_period_str}.csv")


def _read_table(path, kind):
    if not os.path.exists(path):
        raise FileNotFoundError(f"{kind}-frame RC lookup file not found @ {path}")
    with open(path, newline="") as fh:
        rows = [r for r in csv.reader(fh) if r]
    # the reference skips the first column of both rows (lookup.py:107,118), so QP 0 is never in the table
    return {int(float(q)): int(float(v)) for q, v in zip(rows[0][1:], rows[1][1:])}


def get_combined_lookup_table(file_path_i, file_path_p):
    ti, tp = _read_table(file_path_i, "I"), _read_table(file_path_p, "P")
    table = {}
    for qp in list(ti) + [q for q in tp if q not in ti]:
        i_bits, p_bits = ti.get(qp, 0), tp.get(qp, 0)
        table[qp] = {}
        if qp in ti:
            table[qp]["I"] = i_bits
        if qp in tp:
            table[qp]["P"] = p_bits
        table[qp]["C"] = (i_bits + p_bits) // 2
    return table


def generate_rc_lookup(metric_files, params):
    """Average bits per block row for every QP, from metrics.csv files of earlier runs (lookup.py:19-76).
    The QP and block size are taken from the run directory name `<i>_<r>_<qp>_...`."""
    ec = params.encoder_config
    kind = "I" if ec.I_Period == 1 else "P"
    acc = {}
    for path in metric_files:
        m = re.match(r"(\d+)_(-?\d+)(?:\.0)?_(\d+)_", os.path.basename(os.path.dirname(path)))
        if not m:
            raise ValueError(f"cannot parse block size / qp from {path}")
        bs, qp = int(m.group(1)), int(m.group(3))
        rows_per_frame = params.height // bs
        with open(path, newline="") as fh:
            rd = csv.reader(fh)
            next(rd)
            for r in rd:
                is_i, frame_bits = int(r[1]) == 1, int(r[5]) * 8
                if (kind == "I") == is_i:
                    a = acc.setdefault(qp, [0, 0])
                    a[0] += frame_bits
                    a[1] += rows_per_frame
    out = rc_lookup_file_path(ec)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    qps = sorted(acc)
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(qps)
        w.writerow([round(acc[q][0] / acc[q][1]) if acc[q][1] else 0 for q in qps])
    return out
