"""CPU restatement of the reference's rate-controlled frame loop (RCflag 1/2/3) on top of the C oracle's
frame functions.  TEST INFRASTRUCTURE ONLY.

Follows encoder/encoder.py:72-98,174-201 (first / second pass, scene change), encoder/Frame.py:155-188
(get_overage_ratios, get_rc_qp -- including `frame_type` always being 'I'), encoder/PFrame.py:47-49 and
encoder/IFrame.py:35 (prev_frame_avg_qp) and encoder/RateControl/RateControl.py:5-43.  The arithmetic of
every frame is done by oracle/bvc_oracle.c; only the scalar control logic lives here (pure Python, small cases).
"""
from __future__ import annotations

from statistics import mean

import numpy as np

from oracle import bindings as ob


def find_rc_qp_for_row(budget, table, kind="I", scaling=1):
    for qp in sorted(table):
        if table[qp][kind] * scaling <= budget:
            return qp
    return max(table)


class _F:  # the few attributes the control logic passes from frame to frame
    def __init__(self):
        self.rc_qp_per_row, self.bits_per_row, self.intra, self.res = [], [], False, None


def _avg_qp(prev):
    return int(mean(prev.rc_qp_per_row) - 0.1) + 1 if prev.rc_qp_per_row else 0


def _encode(cfg, cur, refs, hps, intra, rcflag, table, budget, prev, first_pass, prev_pass):
    rows = cfg.height // cfg.block
    f = _F()
    f.intra = intra
    avg = _avg_qp(prev)
    state = {"budget": budget, "qp": cfg.qp}

    def cb(row, prev_row_bits):
        state["budget"] -= prev_row_bits                      # self.bit_budget -= row_bits_consumed (PFrame.py:80)
        qp = state["qp"]
        if rcflag == 1:
            qp = find_rc_qp_for_row(state["budget"] / (rows - row), table, "I")
        elif rcflag > 1:
            if first_pass:
                qp = avg
            else:
                share = prev_pass.bits_per_row[row] / sum(prev_pass.bits_per_row)
                qp = find_rc_qp_for_row(budget * share, table, "I", 1)
        state["qp"] = qp
        f.rc_qp_per_row.append(qp)
        return qp

    if intra:
        f.res = ob.encode_iframe(cfg, cur, qp_callback=cb)
    else:
        f.res = ob.encode_pframe(cfg, cur, refs, hps if cfg.frac else None, qp_callback=cb)
    f.bits_per_row = [int(b) for b in f.res.bits_per_row]
    return f


def encode_video_rc(frames: np.ndarray, cfg: ob.Config, rcflag: int, target_br: float, table: dict, frame_rate: int = 30):
    """Returns (container bytes, recon planes, [row QPs per coded frame], [is_intra per frame])."""
    n, H, W = frames.shape
    rows = H // cfg.block
    budget = target_br / frame_rate
    refs, hps = [], []
    prev = _F()
    prev.rc_qp_per_row = [cfg.qp]
    out = bytearray()
    recon = np.empty_like(frames)
    qps, kinds = [], []
    for idx in range(1, n + 1):
        cur = frames[idx - 1]
        intra = (idx - 1) % cfg.i_period == 0
        if intra:
            refs, hps = [], []
        first = _encode(cfg, cur, refs, hps, intra, rcflag, table, budget, prev, True, None)
        fr = first
        if rcflag > 1:
            bits = first.res.coef_nbits + first.res.pred_nbits + 8 * 6
            over_p = bits / (table[cfg.qp]["P"] * rows)       # get_overage_ratios, Frame.py:155-163
            scene = (not first.intra) and over_p > 1.3
            intra2 = scene or first.intra
            if intra2:
                refs, hps = [], []
            fr = _encode(cfg, cur, refs, hps, intra2, rcflag, table, budget, prev, False, first)
        r = fr.res
        pb, cb = (r.pred_nbits + 7) // 8, (r.coef_nbits + 7) // 8
        out += bytes([1 if fr.intra else 0]) + pb.to_bytes(2, "big") + r.pred_bytes + cb.to_bytes(3, "big") + r.coef_bytes
        recon[idx - 1] = r.recon
        qps.append(list(fr.rc_qp_per_row))
        kinds.append(bool(fr.intra))
        refs.append(r.recon)
        hps.append(ob.halfpel_plane(r.recon) if cfg.frac else r.recon)
        if len(refs) > cfg.nref:
            refs.pop(0)
            hps.pop(0)
        prev = fr
    return bytes(out), recon, qps, kinds


def measure_table(frames: np.ndarray, block: int, qps=range(0, 12), nref=1):
    """A lookup table in the reference's format {qp: {'I','P','C'}} measured with the oracle on `frames`
    (average bits per block row, I_Period 1 for 'I' and one long GOP for 'P'; lookup.py:19-76)."""
    n, H, W = frames.shape
    rows = H // block
    table = {}
    lg = block.bit_length() - 1
    for qp in qps:
        if qp > lg + 7 or qp == 0:      # qp 0 never makes it into the parsed table (lookup.py:107,118)
            continue
        vals = {}
        for kind, ip in (("I", 1), ("P", n)):
            cfg = ob.make_config(W, H, block, 2, qp, nref=nref, i_period=ip)
            data, _ = ob.encode_clip(cfg, frames, want_recon=False)
            from tests.golden_util import split_container
            recs = split_container(data)
            sel = [6 + len(p) + len(c) for m, p, c in recs if (m == 1) == (kind == "I")]
            vals[kind] = round(sum(sel) * 8 / (len(sel) * rows))
        vals["C"] = (vals["I"] + vals["P"]) // 2
        table[qp] = vals
    return table
