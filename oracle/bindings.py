"""ctypes bindings of oracle/libbvc_oracle.so (see oracle/bvc_oracle.h).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("width", "height", "block", "range", "qp", "nref", "fastme", "frac", "i_period")]


class Bits(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("nbits", C.c_size_t), ("cap_bytes", C.c_size_t)]


class FrameOut(C.Structure):
    _fields_ = [("recon", C.c_void_p), ("levels", C.c_void_p), ("mv", C.c_void_p), ("sad", C.c_void_p),
                ("modes", C.c_void_p), ("resid_mc", C.c_void_p), ("resid_nomc", C.c_void_p),
                ("pred_bits", Bits), ("coef_bits", Bits), ("bits_per_row", C.c_void_p),
                ("avg_mae", C.c_double), ("mae_comparisons", C.c_int64),
                ("qp_cb", C.c_void_p), ("qp_user", C.c_void_p), ("qp_used", C.c_void_p)]


QP_CALLBACK = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.c_int64)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libbvc_oracle.so")
    src = os.path.join(_HERE, "bvc_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.bvo_full_search_block.restype = C.c_int32
        L.bvo_fast_me_block.restype = C.c_int32
        L.bvo_me_frame.restype = C.c_int64
        L.bvo_dct_ct.restype = C.POINTER(C.c_double)
        L.bvo_dct_w.restype = C.POINTER(C.c_double)
        L.bvo_q_shift.restype = C.c_int
        L.bvo_eg_len.restype = C.c_int
        L.bvo_rle.restype = C.c_int
        L.bvo_encode_clip.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def make_config(width, height, block, search_range, qp, nref=1, fastme=False, frac=False, i_period=1):
    return Config(width, height, block, search_range, qp, nref, int(fastme), int(frac), i_period)


def halfpel_plane(ref: np.ndarray) -> np.ndarray:
    H, W = ref.shape
    out = np.empty((2 * H, 2 * W), dtype=np.uint8)
    lib().bvo_halfpel_plane(_p(np.ascontiguousarray(ref)), W, H, _p(out))
    return out


def _plane_array(planes):
    arr = (C.c_void_p * len(planes))()
    keep = []
    for i, p in enumerate(planes):
        p = np.ascontiguousarray(p, dtype=np.uint8)
        keep.append(p)
        arr[i] = p.ctypes.data
    return arr, keep


def full_search_block(cur: np.ndarray, refs, ox: int, oy: int, bs: int, search_range: int, frac: bool = False):
    """find_lowest_mae_block for one block (block_predictor.py:61-91) -> ((mvx, mvy, ref), SAD).  refs: integer planes, or
    half-pel planes (2H x 2W) when frac."""
    cur = np.ascontiguousarray(cur, dtype=np.uint8)
    H, W = cur.shape
    arr, keep = _plane_array(refs)
    mv = (C.c_int32 * 3)()
    sad = lib().bvo_full_search_block(_p(cur), W, H, int(ox), int(oy), int(bs), arr, len(refs), int(search_range), int(bool(frac)), mv, None)
    return (int(mv[0]), int(mv[1]), int(mv[2])), int(sad)


def me_frame(cfg: Config, cur: np.ndarray, planes):
    """Frame-level motion estimation.  planes: integer planes, or half-pel planes when cfg.frac."""
    nblk = (cfg.width // cfg.block) * (cfg.height // cfg.block)
    mv = np.zeros((nblk, 3), dtype=np.int32)
    sad = np.zeros(nblk, dtype=np.int32)
    arr, keep = _plane_array(planes)
    cur = np.ascontiguousarray(cur, dtype=np.uint8)
    total = lib().bvo_me_frame(C.byref(cfg), _p(cur), arr, len(planes), _p(mv), _p(sad))
    return mv, sad, int(total)


def fdct(res: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(res, dtype=np.int16)
    out = np.empty(b.shape, dtype=np.float64)
    lib().bvo_fdct(_p(b), b.shape[0], _p(out))
    return out


def idct(coef: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(coef, dtype=np.float64)
    out = np.empty(b.shape, dtype=np.float64)
    lib().bvo_idct(_p(b), b.shape[0], _p(out))
    return out


def dct_tables(bs: int):
    L = lib()
    ct = np.ctypeslib.as_array(L.bvo_dct_ct(bs), shape=(bs, bs)).copy()
    w = np.ctypeslib.as_array(L.bvo_dct_w(bs), shape=(bs, bs)).copy()
    return ct, w


def transform_block(res: np.ndarray, pred: np.ndarray, qp: int):
    bs = res.shape[0]
    res = np.ascontiguousarray(res, dtype=np.int16)
    pred = np.ascontiguousarray(pred, dtype=np.int16)
    level = np.empty((bs, bs), dtype=np.int16)
    recon = np.empty((bs, bs), dtype=np.uint8)
    idc = np.empty((bs, bs), dtype=np.float64)
    coef = np.empty((bs, bs), dtype=np.float64)
    lib().bvo_transform_block(_p(res), _p(pred), bs, qp, _p(level), _p(recon), _p(idc), _p(coef))
    return level, recon, idc, coef


def _bits_to_bytes(b: Bits):
    n = (b.nbits + 7) // 8
    return bytes(C.cast(b.data, C.POINTER(C.c_uint8 * n)).contents) if n else b"", int(b.nbits)


class FrameResult:
    pass


def _encode_frame(cfg: Config, cur, refs, hp_refs, qp_rows, intra: bool, qp_callback=None):
    L = lib()
    H, W, bs = cfg.height, cfg.width, cfg.block
    nblk = (W // bs) * (H // bs)
    rows = H // bs
    r = FrameResult()
    r.recon = np.zeros((H, W), dtype=np.uint8)
    r.levels = np.zeros((H, W), dtype=np.int16)
    r.mv = np.zeros((nblk, 3), dtype=np.int32)
    r.sad = np.zeros(nblk, dtype=np.int32)
    r.modes = np.zeros(nblk, dtype=np.int32)
    r.resid_mc = np.zeros((H, W), dtype=np.int8)
    r.resid_nomc = np.zeros((H, W), dtype=np.int8)
    r.bits_per_row = np.zeros(rows, dtype=np.int64)
    fo = FrameOut()
    fo.recon, fo.levels, fo.mv, fo.sad, fo.modes = _p(r.recon), _p(r.levels), _p(r.mv), _p(r.sad), _p(r.modes)
    fo.resid_mc, fo.resid_nomc, fo.bits_per_row = _p(r.resid_mc), _p(r.resid_nomc), _p(r.bits_per_row)
    L.bvo_bits_init(C.byref(fo.pred_bits))
    L.bvo_bits_init(C.byref(fo.coef_bits))
    cur = np.ascontiguousarray(cur, dtype=np.uint8)
    qp = None
    if qp_rows is not None:
        qp = np.ascontiguousarray(qp_rows, dtype=np.int32)
    r.qp_used = np.zeros(rows, dtype=np.int32)
    fo.qp_used = _p(r.qp_used)
    cb = None
    if qp_callback is not None:   # qp_callback(row, prev_row_bits) -> qp
        cb = QP_CALLBACK(lambda user, row, bits: int(qp_callback(int(row), int(bits))))
        fo.qp_cb = C.cast(cb, C.c_void_p)
    if intra:
        L.bvo_encode_iframe(C.byref(cfg), _p(cur), _p(qp) if qp is not None else None, C.byref(fo))
    else:
        ra, k1 = _plane_array(refs)
        ha, k2 = _plane_array(hp_refs if hp_refs is not None else refs)
        L.bvo_encode_pframe(C.byref(cfg), _p(cur), ra, ha, len(refs), _p(qp) if qp is not None else None, C.byref(fo))
    r.pred_bytes, r.pred_nbits = _bits_to_bytes(fo.pred_bits)
    r.coef_bytes, r.coef_nbits = _bits_to_bytes(fo.coef_bits)
    L.bvo_bits_free(C.byref(fo.pred_bits))
    L.bvo_bits_free(C.byref(fo.coef_bits))
    r.avg_mae = float(fo.avg_mae)
    r.mae_comparisons = int(fo.mae_comparisons)
    return r


def encode_pframe(cfg, cur, refs, hp_refs=None, qp_rows=None, qp_callback=None):
    return _encode_frame(cfg, cur, refs, hp_refs, qp_rows, False, qp_callback)


def encode_iframe(cfg, cur, qp_rows=None, qp_callback=None):
    return _encode_frame(cfg, cur, None, None, qp_rows, True, qp_callback)


def encode_clip(cfg: Config, frames: np.ndarray, nthreads: int = 1, want_recon: bool = True):
    """encode_video frame loop + container for RCflag=0.  frames: (n,H,W) uint8, already padded."""
    L = lib()
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n = frames.shape[0]
    out = C.POINTER(C.c_uint8)()
    out_len = C.c_size_t(0)
    recon = np.empty_like(frames) if want_recon else None
    rc = L.bvo_encode_clip(C.byref(cfg), _p(frames), n, 1, C.byref(out), C.byref(out_len),
                           _p(recon) if recon is not None else None, nthreads)
    data = bytes(C.cast(out, C.POINTER(C.c_uint8 * out_len.value)).contents) if out_len.value else b""
    L.bvo_free(out)
    if rc == -2:
        raise OverflowError("payload length does not fit the container field (encoder.py:108-117)")
    if rc != 0:
        raise RuntimeError(f"bvo_encode_clip rc={rc}")
    return data, recon


def decode_clip(cfg: Config, data: bytes, max_frames: int, details: bool = False):
    """decode_video (decoder.py:26-87).  Returns decoded planes (n,H,W) [, levels, pred, qp_rows, kinds]."""
    L = lib()
    H, W, bs = cfg.height, cfg.width, cfg.block
    nblk, rows = (W // bs) * (H // bs), H // bs
    buf = np.frombuffer(data, dtype=np.uint8)
    frames = np.zeros((max_frames, H, W), np.uint8)
    n = C.c_int(0)
    lev = np.zeros((max_frames, H, W), np.int16) if details else None
    pred = np.zeros((max_frames, nblk, 3), np.int32) if details else None
    qps = np.zeros((max_frames, rows), np.int32) if details else None
    kinds = np.zeros(max_frames, np.uint8) if details else None
    L.bvo_decode_clip.restype = C.c_int
    rc = L.bvo_decode_clip(C.byref(cfg), _p(buf), C.c_size_t(buf.size), int(max_frames), _p(frames), C.byref(n),
                           _p(lev) if details else None, _p(pred) if details else None, _p(qps) if details else None,
                           _p(kinds) if details else None)
    if rc != 0:
        raise ValueError("malformed stream")
    k = n.value
    if details:
        return frames[:k], lev[:k], pred[:k], qps[:k], kinds[:k]
    return frames[:k]
