"""Harness that imports the UNMODIFIED Python reference from /root/reference (read-only) so the C
oracle can be validated against it and golden vectors can be generated from it.

TEST INFRASTRUCTURE ONLY.  It works only in the build container (where /root/reference exists);
nothing that runs on the GPU box imports this module -- the goldens it produces are committed under
tests/golden/ by oracle/gen_golden.py.

What is patched (by monkeypatch, never by editing):
  * import shims for packages missing from this image: bitarray, skimage.metrics, matplotlib
    (oracle/shims/), import-time only except the bit container whose semantics are restated;
  * `encoder.IFrame.mean` guarded against an empty list (reference crash Q13, IFrame.py:35,74 --
    the value is unused when RCflag <= 1);
  * DCT mode (the reference runs scipy.fftpack in float32, encoder/dct.py:9-18):
        "asis"         -- untouched float32 SciPy
        "fp64_scipy"   -- same SciPy call on float64 input
        "fp64_defined" -- the defined fp64 transform of DESIGN.md, executed by the C oracle through
                          ctypes (bit-exact contract for levels / recon / bitstream)
  * `encoder.encoder.__file__` redirected to a scratch dir so the `results.csv` side effect
    (encoder.py:167-169) does not try to write into the read-only reference tree;
  * for resolutions without an RC lookup CSV (anything but CIF/QCIF x {8,16}), a constant lookup
    table (values only feed a log line when RCflag == 0, Frame.py:158-163).
"""
from __future__ import annotations

import ctypes
import importlib
import logging
import os
import statistics
import sys
import tempfile
from collections import deque

import numpy as np

REFERENCE_ROOT = os.environ.get("BVC_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")

_state = {"loaded": False, "dct_mode": None}


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "encoder"))


def _oracle_lib():
    from oracle import bindings  # local import: keeps this module importable without the .so
    return bindings.lib()


def load_reference():
    """Import the reference package tree (once) and return a namespace of its modules."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    if not _state["loaded"]:
        for p in (REFERENCE_ROOT, _SHIMS):
            if p in sys.path:
                sys.path.remove(p)
        # appended, not prepended: the reference has a top-level `tests` package that must not shadow ours
        # (its hot-path modules -- encoder, common, decoder, file_io, ... -- have no namesakes here)
        sys.path.append(_SHIMS)
        sys.path.append(REFERENCE_ROOT)
        # the reference's `tests` package name collides with ours only if imported; we never do.
        importlib.import_module("encoder.encoder")
        importlib.import_module("decoder")
        logging.getLogger().setLevel(logging.WARNING)  # reference sets the ROOT logger to INFO
        import encoder.IFrame as ri
        ri.mean = lambda xs: statistics.mean(xs) if len(xs) else 0  # Q13 guard
        import encoder.encoder as re_
        scratch = tempfile.mkdtemp(prefix="bvc_ref_results_")
        os.makedirs(os.path.join(scratch, "encoder"), exist_ok=True)
        re_.__file__ = os.path.join(scratch, "encoder", "encoder.py")
        _state["loaded"] = True
        _state["orig_dct"] = None
    return _modules()


class _NS:
    pass


def _modules():
    ns = _NS()
    import encoder.encoder as m_encoder
    import encoder.dct as m_dct
    import encoder.Frame as m_frame
    import encoder.PFrame as m_pframe
    import encoder.IFrame as m_iframe
    import encoder.block_predictor as m_bp
    import encoder.entropy_encoder as m_ee
    import encoder.params as m_params
    import input_parameters as m_ip
    import decoder as m_decoder
    import common as m_common
    ns.encoder, ns.dct, ns.Frame, ns.PFrame, ns.IFrame = m_encoder, m_dct, m_frame, m_pframe, m_iframe
    ns.block_predictor, ns.entropy_encoder, ns.params = m_bp, m_ee, m_params
    ns.input_parameters, ns.decoder, ns.common = m_ip, m_decoder, m_common
    return ns


def set_dct_mode(mode: str):
    """Select the DCT arithmetic used by the imported reference (see module docstring)."""
    ns = load_reference()
    if _state.get("orig_dct") is None:
        _state["orig_dct"] = (ns.dct.apply_dct_2d, ns.dct.apply_idct_2d)
    if mode == "asis":
        f, g = _state["orig_dct"]
    elif mode == "fp64_scipy":
        from scipy.fftpack import dct, idct

        def f(block):
            block = np.asarray(block).astype(np.float64)
            return dct(dct(block.T, norm="ortho").T, norm="ortho")

        def g(block):
            block = np.asarray(block).astype(np.float64)
            return idct(idct(block.T, norm="ortho").T, norm="ortho")
    elif mode == "fp64_defined":
        lib = _oracle_lib()

        def f(block):
            b = np.ascontiguousarray(block, dtype=np.int16)
            n = b.shape[0]
            out = np.empty((n, n), dtype=np.float64)
            lib.bvo_fdct(b.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p))
            return out

        def g(block):
            b = np.ascontiguousarray(block, dtype=np.float64)
            n = b.shape[0]
            out = np.empty((n, n), dtype=np.float64)
            lib.bvo_idct(b.ctypes.data_as(ctypes.c_void_p), n, out.ctypes.data_as(ctypes.c_void_p))
            return out
    else:
        raise ValueError(mode)
    # the names are from-imported into four namespaces (Frame.py:12, PFrame.py:14, IFrame.py:9)
    for mod in (ns.dct, ns.Frame, ns.PFrame, ns.IFrame):
        if hasattr(mod, "apply_dct_2d"):
            mod.apply_dct_2d = f
        if hasattr(mod, "apply_idct_2d"):
            mod.apply_idct_2d = g
    _state["dct_mode"] = mode


def make_config(ns, *, block, search_range, qp, i_period, nref=1, fastme=False, frac=False,
                width=352, height=288):
    return ns.params.EncoderConfig(block, search_range, i_period, qp, nRefFrames=nref, fastME=fastme,
                                   fracMeEnabled=frac, RCflag=0, targetBR=0, resolution=(width, height))


def ref_encode_video(frames: np.ndarray, *, block, search_range, qp, i_period, nref=1, fastme=False,
                     frac=False, dct_mode="fp64_defined"):
    """Run the reference's own encode_video() on `frames` (n,H,W uint8) through a scratch file and
    return its outputs: container bytes, reconstructed planes, level planes, mv.txt text."""
    ns = load_reference()
    set_dct_mode(dct_mode)
    n, H, W = frames.shape
    ec = make_config(ns, block=block, search_range=search_range, qp=qp, i_period=i_period, nref=nref,
                     fastme=fastme, frac=frac, width=W, height=H)
    lookup_ok = os.path.exists(ns.encoder.rc_lookup_file_path(ec, "I"))
    orig_lookup = ns.encoder.get_combined_lookup_table
    if not lookup_ok:
        ns.encoder.get_combined_lookup_table = lambda a, b: {q: {"I": 1000, "P": 1000, "C": 1000} for q in range(0, 13)}
    try:
        with tempfile.TemporaryDirectory(prefix="bvc_ref_") as td:
            yfile = os.path.join(td, "clip.y")
            with open(yfile, "wb") as fh:
                fh.write(np.ascontiguousarray(frames, dtype=np.uint8).tobytes())
            params = ns.input_parameters.InputParameters(yfile, W, H, ec, frames_to_process=n)
            ns.encoder.encode_video(params)
            fio = ns.encoder.FileIOHelper(params)
            out = {
                "encoded": open(fio.get_encoded_file_name(), "rb").read(),
                "recon": np.fromfile(fio.get_mc_reconstructed_file_name(), dtype=np.uint8).reshape(n, H, W),
                "levels": np.fromfile(fio.get_quant_dct_coff_fh_file_name(), dtype=np.int16).reshape(n, H, W),
                "resid_mc": np.fromfile(fio.get_residual_w_mc_file_name(), dtype=np.uint8).reshape(n, H, W),
                "resid_nomc": np.fromfile(fio.get_residual_wo_mc_file_name(), dtype=np.uint8).reshape(n, H, W),
                "mv_txt": open(fio.get_mv_file_name(), "rt").read(),
                "metrics_csv": open(fio.get_metrics_csv_file_name(), "rt").read(),
            }
    finally:
        ns.encoder.get_combined_lookup_table = orig_lookup
    return out


def ref_encode_frames(frames: np.ndarray, *, block, search_range, qp, i_period, nref=1, fastme=False,
                      frac=False, dct_mode="fp64_defined"):
    """Drive IFrame/PFrame objects exactly like the reference frame loop (encoder.py:75-155) but
    in memory, returning the per-frame objects (for MV / mode / SAD / bit-string level checks)."""
    ns = load_reference()
    set_dct_mode(dct_mode)
    n, H, W = frames.shape
    ec = make_config(ns, block=block, search_range=search_range, qp=qp, i_period=i_period, nref=nref,
                     fastme=fastme, frac=frac, width=W, height=H)
    refs = deque(maxlen=nref)
    irefs = deque(maxlen=nref)
    prev = ns.Frame.Frame()
    prev.rc_qp_per_row = [qp]
    out = []
    for idx in range(1, n + 1):
        cur = ns.common.pad_frame(frames[idx - 1], block)
        if (idx - 1) % i_period == 0:
            fr = ns.IFrame.IFrame(cur)
            refs.clear()
            irefs.clear()
        else:
            fr = ns.PFrame.PFrame(cur, refs, irefs)
        fr.is_first_pass = True
        fr.prev_frame = prev
        fr.index = idx
        fr.bit_budget = 0
        fr.encode_mc_q_dct(ec)
        out.append(fr)
        refs.append(fr.reconstructed_frame)
        irefs.append(ns.block_predictor.build_pre_interpolated_buffer(fr.reconstructed_frame)
                     if frac else np.zeros((2, 2), dtype=np.uint8))
        prev = fr
    return out
