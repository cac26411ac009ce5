"""ctypes binding of libbvc_b200.so (include/bvc.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BVC_OK, BVC_ERR_INVALID, BVC_ERR_CUDA, BVC_ERR_OVERFLOW, BVC_ERR_NOMEM, BVC_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5

EXPORTS = [
    "bvc_create", "bvc_destroy", "bvc_last_error", "bvc_set_qp", "bvc_encode_iframe", "bvc_encode_pframe",
    "bvc_frame_begin", "bvc_frame_encode_row", "bvc_frame_end", "bvc_me_search", "bvc_interp_halfpel", "bvc_dct_quant_recon", "bvc_encode_clip", "bvc_clip_upload",
    "bvc_encode_clip_resident", "bvc_launch_count", "bvc_last_kernel_times", "bvc_me_work_per_frame", "bvc_set_lane_groups", "bvc_decode_clip", "bvc_decode_frame", "bvc_clip_upload_i420", "bvc_set_fastme_direct",
    "bvc_encode_clip_device", "bvc_container_download", "bvc_host_register", "bvc_host_unregister", "bvc_measure_peaks", "bvc_set_rate_control", "bvc_set_stream_slot_bytes",
]


class BvcError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "height", "block_size", "search_range", "qp", "nref_frames",
                                       "fast_me", "frac_me", "i_period")]


class FrameOut(C.Structure):
    _fields_ = [("recon", C.c_void_p), ("levels", C.c_void_p), ("mv", C.c_void_p), ("sad", C.c_void_p),
                ("modes", C.c_void_p), ("resid_mc", C.c_void_p), ("resid_nomc", C.c_void_p),
                ("pred_bytes", C.c_void_p), ("pred_cap", C.c_size_t), ("coef_bytes", C.c_void_p),
                ("coef_cap", C.c_size_t), ("bits_per_row", C.c_void_p),
                ("pred_nbits", C.c_int64), ("coef_nbits", C.c_int64), ("avg_mae", C.c_double),
                ("mae_comparisons", C.c_int64)]


def library_path() -> str:
    return os.path.join(_HERE, "libbvc_b200.so")


def load_library():
    """Load libbvc_b200.so.  No fallback: a missing library is an error, not a slow path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise BvcError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       f"or `make -C basic_video_codec_b200/csrc` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    L.bvc_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Params), C.c_int]
    L.bvc_create.restype = C.c_int
    L.bvc_destroy.argtypes = [C.c_void_p]
    L.bvc_destroy.restype = None
    L.bvc_last_error.argtypes = [C.c_void_p]
    L.bvc_last_error.restype = C.c_char_p
    L.bvc_set_qp.argtypes = [C.c_void_p, C.c_int]
    L.bvc_encode_iframe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(FrameOut)]
    L.bvc_encode_pframe.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.POINTER(FrameOut)]
    L.bvc_frame_begin.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int]
    L.bvc_frame_encode_row.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    L.bvc_frame_end.argtypes = [C.c_void_p, C.POINTER(FrameOut)]
    L.bvc_me_search.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_int64)]
    L.bvc_interp_halfpel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.bvc_dct_quant_recon.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
    L.bvc_encode_clip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
    L.bvc_clip_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.bvc_encode_clip_resident.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
    L.bvc_set_lane_groups.argtypes = [C.c_void_p, C.c_int]
    L.bvc_set_fastme_direct.argtypes = [C.c_void_p, C.c_int]
    L.bvc_clip_upload_i420.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.bvc_decode_clip.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
    L.bvc_decode_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bvc_launch_count.argtypes = [C.c_void_p]
    L.bvc_launch_count.restype = C.c_int64
    L.bvc_last_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.bvc_me_work_per_frame.argtypes = [C.c_void_p, C.c_int]
    L.bvc_me_work_per_frame.restype = C.c_int64
    L.bvc_encode_clip_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.POINTER(C.c_size_t)]
    L.bvc_container_download.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    L.bvc_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.bvc_host_unregister.argtypes = [C.c_void_p]
    L.bvc_set_rate_control.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    L.bvc_set_stream_slot_bytes.argtypes = [C.c_void_p, C.c_size_t]
    L.bvc_measure_peaks.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    _LIB = L
    return L


def _raise(code: int, msg: str):
    """Map C status codes onto the exception types the reference raises."""
    if code == BVC_ERR_INVALID:
        raise ValueError(msg)
    if code == BVC_ERR_OVERFLOW:
        raise OverflowError(msg)
    if code == BVC_ERR_NOMEM:
        raise MemoryError(msg)
    if code == BVC_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise BvcError(msg)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class FrameResult:
    """Plain record of one encoded frame (numpy arrays + bit strings)."""
    __slots__ = ("recon", "levels", "mv", "sad", "modes", "resid_mc", "resid_nomc", "pred_bytes", "pred_nbits",
                 "coef_bytes", "coef_nbits", "bits_per_row", "avg_mae", "mae_comparisons")


class Context:
    """One encoder context = one GPU + one geometry/parameter set (bvc_ctx)."""

    def __init__(self, width, height, block_size, search_range, qp, nref_frames=1, fast_me=False, frac_me=False,
                 i_period=1, device=0, max_lanes=1):
        self._L = load_library()
        self._h = C.c_void_p()
        self.params = Params(int(width), int(height), int(block_size), int(max(search_range, 0)), int(qp),
                             int(nref_frames), int(bool(fast_me)), int(bool(frac_me)), int(i_period))
        rc = self._L.bvc_create(C.byref(self._h), int(device), C.byref(self.params), int(max_lanes))
        if rc != BVC_OK:
            msg = self._L.bvc_last_error(None).decode()
            self._h = C.c_void_p()
            _raise(rc, msg)
        self.W, self.H, self.bs = int(width), int(height), int(block_size)
        self.nblk = (self.W // self.bs) * (self.H // self.bs)
        self.rows = self.H // self.bs
        self.max_lanes = int(max_lanes)
        self.lane_groups = int(os.environ.get("BVC_LANE_GROUPS", "2"))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.bvc_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != BVC_OK:
            _raise(rc, self._L.bvc_last_error(self._h).decode())

    # ---- frame level -------------------------------------------------------------------------
    def _alloc_out(self, debug_planes=True):
        H, W = self.H, self.W
        r = FrameResult()
        r.recon = np.empty((H, W), np.uint8)
        r.levels = np.empty((H, W), np.int16)
        r.mv = np.zeros((self.nblk, 3), np.int32)
        r.sad = np.empty(self.nblk, np.int32)
        r.modes = np.zeros(self.nblk, np.int32)
        r.resid_mc = np.empty((H, W), np.int8) if debug_planes else None
        r.resid_nomc = np.zeros((H, W), np.int8) if debug_planes else None
        r.bits_per_row = np.empty(self.rows, np.int64)
        coef_cap = self.nblk * (832 if self.bs == 16 else 192 if self.bs == 8 else 48) + 64
        pred_cap = self.nblk * 12 + self.rows * 8 + 64
        pred = np.empty(pred_cap, np.uint8)
        coef = np.empty(coef_cap, np.uint8)
        fo = FrameOut()
        fo.recon, fo.levels, fo.mv, fo.sad, fo.modes = _p(r.recon), _p(r.levels), _p(r.mv), _p(r.sad), _p(r.modes)
        fo.resid_mc, fo.resid_nomc = _p(r.resid_mc), _p(r.resid_nomc)
        fo.pred_bytes, fo.pred_cap, fo.coef_bytes, fo.coef_cap = _p(pred), pred_cap, _p(coef), coef_cap
        fo.bits_per_row = _p(r.bits_per_row)
        return r, fo, pred, coef

    @staticmethod
    def _finish_out(r, fo, pred, coef):
        r.pred_nbits, r.coef_nbits = int(fo.pred_nbits), int(fo.coef_nbits)
        r.pred_bytes = pred[: (r.pred_nbits + 7) // 8].tobytes()
        r.coef_bytes = coef[: (r.coef_nbits + 7) // 8].tobytes()
        r.avg_mae, r.mae_comparisons = float(fo.avg_mae), int(fo.mae_comparisons)
        return r

    def _check_frame(self, cur):
        cur = np.ascontiguousarray(cur, dtype=np.uint8)
        if cur.shape != (self.H, self.W):
            raise ValueError(f"frame shape {cur.shape} != {(self.H, self.W)}")
        return cur

    def _frame(self, cur, refs, qp_rows, intra, debug_planes=True):
        cur = self._check_frame(cur)
        r, fo, pred, coef = self._alloc_out(debug_planes)
        qp = np.ascontiguousarray(qp_rows, dtype=np.int32) if qp_rows is not None else None
        if intra:
            self._check(self._L.bvc_encode_iframe(self._h, _p(cur), _p(qp), C.byref(fo)))
        else:
            keep = [self._check_frame(x) for x in refs]
            arr = (C.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
            self._check(self._L.bvc_encode_pframe(self._h, _p(cur), arr, len(keep), _p(qp), C.byref(fo)))
        return self._finish_out(r, fo, pred, coef)

    # ---- row level (rate-control feedback loop) ---------------------------------------------------------
    def frame_begin(self, cur, refs=None):
        """Upload a frame (and, for a P frame, its reference window) and run motion estimation."""
        cur = self._check_frame(cur)
        if refs is None:
            self._check(self._L.bvc_frame_begin(self._h, _p(cur), None, 0, 1))
        else:
            keep = [self._check_frame(x) for x in refs]
            arr = (C.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
            self._check(self._L.bvc_frame_begin(self._h, _p(cur), arr, len(keep), 0))

    def frame_encode_row(self, row, qp):
        """Encode block row `row` with `qp`; returns the bits it added to both streams."""
        bits = C.c_int64(0)
        self._check(self._L.bvc_frame_encode_row(self._h, int(row), int(qp), C.byref(bits)))
        return int(bits.value)

    def frame_end(self, debug_planes=True):
        r, fo, pred, coef = self._alloc_out(debug_planes)
        self._check(self._L.bvc_frame_end(self._h, C.byref(fo)))
        return self._finish_out(r, fo, pred, coef)

    def encode_iframe(self, cur, qp_rows=None, debug_planes=True):
        return self._frame(cur, None, qp_rows, True, debug_planes)

    def encode_pframe(self, cur, refs, qp_rows=None, debug_planes=True):
        return self._frame(cur, list(refs), qp_rows, False, debug_planes)

    # ---- hooks ---------------------------------------------------------------------------------
    def me_search(self, cur, refs):
        cur = self._check_frame(cur)
        keep = [self._check_frame(x) for x in refs]
        arr = (C.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
        mv = np.zeros((self.nblk, 3), np.int32)
        sad = np.zeros(self.nblk, np.int32)
        cmp_ = C.c_int64(0)
        self._check(self._L.bvc_me_search(self._h, _p(cur), arr, len(keep), _p(mv), _p(sad), C.byref(cmp_)))
        return mv, sad, int(cmp_.value)

    def interp_halfpel(self, ref):
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        out = np.empty((2 * self.H, 2 * self.W), np.uint8)
        self._check(self._L.bvc_interp_halfpel(self._h, _p(ref), _p(out)))
        return out

    # ---- clip level ------------------------------------------------------------------------------
    def encode_clip(self, frames, want_recon=False, out_capacity=None):
        """Host-buffer clip encode (the public end-to-end call).  Returns (container bytes, recon | None)."""
        frames = self._check_clip(frames)
        n = frames.shape[0]
        cap = int(out_capacity or (n * self.W * self.H // 2 + (1 << 20)))
        recon = np.empty_like(frames) if want_recon else None
        for _ in range(6):
            out = np.empty(cap, np.uint8)
            ln = C.c_size_t(0)
            rc = self._L.bvc_encode_clip(self._h, _p(frames), n, _p(out), out.size, C.byref(ln), _p(recon))
            # noisy content at a low QP can need more than the defaults (4 bits per pixel of host buffer, 6 of device slot):
            # the library says which one was too small
            if rc == BVC_ERR_NOMEM and b"device slot" in self._L.bvc_last_error(self._h):
                self._check(self._L.bvc_set_stream_slot_bytes(self._h, C.c_size_t(-1).value))
                continue
            if rc == BVC_ERR_NOMEM and b"staging" in self._L.bvc_last_error(self._h):
                continue            # the library has enlarged it
            if rc == BVC_ERR_NOMEM and out_capacity is None and int(ln.value) > cap:
                cap = max(int(ln.value), 2 * cap)
                continue
            break
        self._check(rc)
        return out[:int(ln.value)].tobytes(), recon

    def _check_clip(self, frames):
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        if frames.ndim != 3 or frames.shape[1:] != (self.H, self.W):
            raise ValueError(f"clip shape {frames.shape} != (n, {self.H}, {self.W})")
        return frames

    @staticmethod
    def _check_buffer(buf, nbytes, what):
        if buf is None:
            return
        if not (isinstance(buf, np.ndarray) and buf.dtype == np.uint8 and buf.flags.c_contiguous and buf.flags.writeable):
            raise ValueError(f"{what} must be a writable C-contiguous uint8 array")
        if buf.size < nbytes:
            raise ValueError(f"{what} holds {buf.size} bytes, {nbytes} needed")

    def encode_clip_into(self, frames, out, recon=None):
        """Same call without Python-side copies: `out` is a caller-owned uint8 buffer (pinned for best
        transfer speed); returns the number of container bytes written."""
        frames = self._check_clip(frames)
        self._check_buffer(out, 6, "out")
        self._check_buffer(recon, frames.size, "recon")
        ln = C.c_size_t(0)
        self._check(self._repeat_if_enlarged(lambda: self._L.bvc_encode_clip(self._h, _p(frames), frames.shape[0], _p(out), out.size,
                                                                             C.byref(ln), _p(recon))))
        return int(ln.value)

    def _repeat_if_enlarged(self, call):
        """Content that codes more bits than the library's staging buffers were sized for makes a clip call fail with
        BVC_ERR_NOMEM after the library has enlarged the buffer in question ("repeat the call"): do that, a few times at most
        (the two reservations can each be found too small once)."""
        rc = call()
        for _ in range(3):
            if rc != BVC_ERR_NOMEM:
                break
            msg = self._L.bvc_last_error(self._h)
            if b"staging" in msg:
                rc = call()
            else:
                break
        return rc

    # ---- decoder ---------------------------------------------------------------------------------
    def decode_clip(self, data, max_frames, details=False, out=None):
        """decode_video on container bytes.  Returns the decoded planes (n,H,W); with details=True also
        (levels, pred (n,nblk,3), qp_rows (n,rows), kinds)."""
        buf = np.frombuffer(bytes(data) if not isinstance(data, np.ndarray) else data, dtype=np.uint8)
        frames = out if out is not None else np.empty((max_frames, self.H, self.W), np.uint8)
        n = C.c_int(0)
        lev = np.zeros((max_frames, self.H, self.W), np.int16) if details else None
        pred = np.zeros((max_frames, self.nblk, 3), np.int32) if details else None
        qps = np.zeros((max_frames, self.rows), np.int32) if details else None
        kinds = np.zeros(max_frames, np.uint8) if details else None
        self._check(self._L.bvc_decode_clip(self._h, _p(buf), buf.size, int(max_frames), _p(frames), C.byref(n), _p(lev), _p(pred),
                                            _p(qps), _p(kinds)))
        k = n.value
        if details:
            return frames[:k], lev[:k], pred[:k], qps[:k], kinds[:k]
        return frames[:k]

    def decode_prediction_data(self, intra, pred_bytes):
        """entropy_decode_prediction_data on its own -> (pred (nblk,3), qp_rows)."""
        pb = np.frombuffer(bytes(pred_bytes), dtype=np.uint8)
        pred = np.zeros((self.nblk, 3), np.int32)
        qps = np.zeros(self.rows, np.int32)
        self._check(self._L.bvc_decode_frame(self._h, int(bool(intra)), _p(pb), pb.size, None, 0, None, 0, None, None, _p(pred), _p(qps)))
        return pred, qps

    def decode_frame(self, intra, pred_bytes, coef_bytes, refs=None):
        """One container record -> (recon, levels, pred (nblk,3), qp_rows)."""
        pb = np.frombuffer(bytes(pred_bytes), dtype=np.uint8)
        cb = np.frombuffer(bytes(coef_bytes), dtype=np.uint8)
        refs = [np.ascontiguousarray(r, dtype=np.uint8) for r in (refs or [])]
        arr = (C.c_void_p * max(1, len(refs)))(*[r.ctypes.data for r in refs])
        recon = np.empty((self.H, self.W), np.uint8)
        lev = np.empty((self.H, self.W), np.int16)
        pred = np.zeros((self.nblk, 3), np.int32)
        qps = np.zeros(self.rows, np.int32)
        self._check(self._L.bvc_decode_frame(self._h, int(bool(intra)), _p(pb), pb.size, _p(cb), cb.size, arr, len(refs), _p(recon),
                                             _p(lev), _p(pred), _p(qps)))
        return recon, lev, pred, qps

    def clip_upload(self, frames):
        frames = self._check_clip(frames)
        self._check(self._L.bvc_clip_upload(self._h, _p(frames), frames.shape[0]))

    def clip_upload_i420(self, yuv, src_w, src_h, nframes):
        """Upload the luma planes of an I420 buffer (padded to the context size with 128 on the device)."""
        yuv = np.ascontiguousarray(np.frombuffer(yuv, dtype=np.uint8) if not isinstance(yuv, np.ndarray) else yuv, dtype=np.uint8)
        need = nframes * (src_w * src_h + 2 * (src_w // 2) * (src_h // 2))
        if yuv.size < need:
            raise ValueError(f"I420 buffer holds {yuv.size} bytes, {need} needed for {nframes} frames")
        self._check(self._L.bvc_clip_upload_i420(self._h, _p(yuv), int(src_w), int(src_h), int(nframes)))

    def encode_clip_resident(self, nframes, out=None):
        if out is None:
            out = np.empty(nframes * self.W * self.H // 2 + (1 << 20), np.uint8)
        self._check_buffer(out, 6, "out")
        ln = C.c_size_t(0)
        self._check(self._repeat_if_enlarged(lambda: self._L.bvc_encode_clip_resident(self._h, int(nframes), _p(out), out.size, C.byref(ln), None)))
        return out, int(ln.value)

    def set_rate_control(self, rc_flag, frame_bit_budget=0.0, table=None):
        """RCflag 1 on the clip path (bvc_set_rate_control).  table: the reference's lookup {qp: {"I": bits per row, ...}}
        (encoder/RateControl/lookup.py:97-118); rc_flag 0 turns rate control off."""
        if not rc_flag:
            self._check(self._L.bvc_set_rate_control(self._h, 0, 0.0, 0, None, None))
            return
        qps = np.array(sorted(q for q in table if "I" in table[q]), dtype=np.int32)
        bits = np.array([int(table[int(q)]["I"]) for q in qps], dtype=np.int64)
        self._check(self._L.bvc_set_rate_control(self._h, int(rc_flag), float(frame_bit_budget), int(qps.size), _p(qps), _p(bits)))

    def set_stream_slot_bytes(self, nbytes):
        """Device bytes reserved per frame for its coefficient stream (bvc_set_stream_slot_bytes; 0 = default)."""
        self._check(self._L.bvc_set_stream_slot_bytes(self._h, int(nbytes)))

    # ---- sharded jobs: container left on the device, fetched into a caller-chosen place ----------------
    def encode_clip_device(self, frames, nframes=None, cap_hint=0):
        """Encode `frames` (host array) or, with frames=None, the first `nframes` resident frames; the container stays in
        device memory.  Returns its length; fetch it with container_download()."""
        ln = C.c_size_t(0)
        if frames is not None:
            frames = self._check_clip(frames)
            nframes = frames.shape[0]
        for _ in range(2):
            rc = self._repeat_if_enlarged(lambda: self._L.bvc_encode_clip_device(self._h, _p(frames), int(nframes), int(cap_hint), C.byref(ln)))
            if rc == BVC_ERR_NOMEM and int(ln.value) > cap_hint:
                cap_hint = int(ln.value)
                continue
            break
        self._check(rc)
        return int(ln.value)

    def container_download(self, dst, dst_offset=0, offset=0, length=None):
        """Copy bytes [offset, offset+length) of the last container into the uint8 array `dst` at dst_offset."""
        if length is None:
            length = dst.size - dst_offset
        self._check_buffer(dst, dst_offset + length, "dst")
        self._check(self._L.bvc_container_download(self._h, C.c_void_p(dst.ctypes.data + int(dst_offset)), int(offset), int(length)))

    # ---- instrumentation -----------------------------------------------------------------------
    def set_lane_groups(self, groups: int):
        """Lane groups of the clip path (bvc_set_lane_groups): 1 = serial, default 2."""
        self._check(self._L.bvc_set_lane_groups(self._h, int(groups)))
        self.lane_groups = int(groups)

    def set_fastme_direct(self, on):
        """FastME evaluation (bvc_set_fastme_direct): 0 / False (default) = automatic (window walk with many frames in
        flight, else transfer tables), 1 / True = every candidate evaluated directly, 2 = SAD map + serial walk,
        3 = window walk (TMA-staged windows, no SAD map), 4 = SAD map + transfer tables.  The output does not depend on it."""
        self._check(self._L.bvc_set_fastme_direct(self._h, int(on)))

    def launch_count(self):
        return int(self._L.bvc_launch_count(self._h))

    KERNEL_CLASSES = ("me", "tq_p", "tq_i", "pack", "halfpel")

    def last_kernel_times(self):
        """{class: (ms, launches)} of the last clip call plus the whole-call device time in ms."""
        ms = (C.c_double * 5)()
        n = (C.c_int64 * 5)()
        clip = C.c_double(0)
        self._L.bvc_last_kernel_times(self._h, ms, n, C.byref(clip))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNEL_CLASSES)}, float(clip.value)

    def me_work_per_frame(self, nref_avail=1):
        return int(self._L.bvc_me_work_per_frame(self._h, int(nref_avail)))


def dct_quant_recon(residual, pred, qp, device=0):
    """apply_dct_and_quantization + reconstruct_block on a batch of blocks (n, bs, bs)."""
    L = load_library()
    residual = np.ascontiguousarray(residual, dtype=np.int16)
    pred = np.ascontiguousarray(pred, dtype=np.int16)
    n, bs, _ = residual.shape
    level = np.empty((n, bs, bs), np.int16)
    recon = np.empty((n, bs, bs), np.uint8)
    idct = np.empty((n, bs, bs), np.float64)
    coef = np.empty((n, bs, bs), np.float64)
    rc = L.bvc_dct_quant_recon(int(device), _p(residual), _p(pred), n, bs, int(qp), _p(level), _p(recon), _p(idct), _p(coef))
    if rc != BVC_OK:
        _raise(rc, L.bvc_last_error(None).decode())
    return level, recon, idct, coef


def host_register(arr):
    """Page-lock the memory of a numpy array the caller keeps alive (bvc_host_register)."""
    L = load_library()
    rc = L.bvc_host_register(C.c_void_p(arr.ctypes.data), arr.nbytes)
    if rc != BVC_OK:
        _raise(rc, L.bvc_last_error(None).decode())


def host_unregister(arr):
    L = load_library()
    L.bvc_host_unregister(C.c_void_p(arr.ctypes.data))


def measure_peaks(device=0):
    """Issue-rate ceilings measured on `device` right now: {"vabsdiff4_thread_ops_per_s", "px_absdiff_per_s",
    "dfma_thread_ops_per_s", "sm_count"} (bvc_measure_peaks)."""
    L = load_library()
    v, f, n = C.c_double(0), C.c_double(0), C.c_int(0)
    rc = L.bvc_measure_peaks(int(device), C.byref(v), C.byref(f), C.byref(n))
    if rc != BVC_OK:
        _raise(rc, "bvc_measure_peaks failed (no usable CUDA device)")
    return {"vabsdiff4_thread_ops_per_s": v.value, "px_absdiff_per_s": 4.0 * v.value, "dfma_thread_ops_per_s": f.value,
            "sm_count": n.value}
