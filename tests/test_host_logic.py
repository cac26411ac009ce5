"""CPU-only checks of the host layer: the C-ABI library loads and exports every declared symbol, the
parameter objects validate like the reference's, and the product never reaches into oracle/."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import basic_video_codec_b200._lib as l
    L = l.load_library()
    hdr = open(os.path.join(ROOT, "include", "bvc.h")).read()
    declared = set(re.findall(r"\b(bvc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(l.EXPORTS)
    for s in declared:
        assert hasattr(L, s), s


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import basic_video_codec_b200 as bvc
    with pytest.raises(bvc.BvcError):
        bvc.Context(64, 64, 16, 4, 3)


def test_encoder_config_validation():
    from basic_video_codec_b200 import EncoderConfig
    with pytest.raises(ValueError):
        EncoderConfig(8, 4, 8, 11)            # qp > log2(8) + 7
    with pytest.raises(ValueError):
        EncoderConfig(8, 4, 8, 3, RCflag=1)   # rate control without a target bitrate
    assert EncoderConfig(16, 16, 8, 3, fastME=True).search_range == -1
    ec = EncoderConfig(8, 4, 8, 3)
    assert (ec.nRefFrames, ec.fastME, ec.fracMeEnabled, ec.RCflag, ec.frame_rate, ec.resolution) == (1, False, False, 0, 30, (352, 288))


def test_product_does_not_touch_oracle():
    """oracle/ is test infrastructure: nothing in the product may import, include, link or load it."""
    pkg = os.path.join(ROOT, "basic_video_codec_b200")
    bad = re.compile(r'(#include\s+["<][^">]*oracle)|(^\s*import\s+oracle)|(^\s*from\s+oracle)|(libbvc_oracle)|(-lbvc_oracle)', re.M)
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                src = open(os.path.join(dp, f)).read()
                assert not bad.search(src), f


def test_bitstring_protocol():
    from basic_video_codec_b200.encoder.Frame import BitString
    b = BitString(bytes([0b10100000]), 3)
    assert len(b) == 3 and b.tobytes() == bytes([0b10100000]) and b.to01() == "101" and bool(b)
    assert not BitString()


def test_dct_tables_agree_with_oracle():
    """The product's generated constant tables (exact decimal arithmetic) and the oracle's independently
    derived ones (long-double libm) must be bit-identical."""
    from oracle import bindings as ob
    hdr = open(os.path.join(ROOT, "basic_video_codec_b200", "csrc", "bvc_dct_tables.h")).read()
    for bs in (2, 4, 8, 16, 32):
        m = re.search(rf"#define BVC_CT{bs}_INIT \{{(.*?)\n\}}", hdr, re.S)
        vals = [float.fromhex(t) for t in re.findall(r"-?0x[0-9a-f.]+p[+-]\d+", m.group(1))]
        ct, w = ob.dct_tables(bs)
        assert len(vals) == bs * bs
        assert [v.hex() for v in vals] == [float(x).hex() for x in ct.ravel()]
        mw = re.search(rf"#define BVC_W{bs}_INIT \{{(.*?)\}}", hdr)
        wv = [float.fromhex(t) for t in re.findall(r"-?0x[0-9a-f.]+p[+-]\d+", mw.group(1))]
        assert set(float(x).hex() for x in w.ravel()) <= set(v.hex() for v in wv)


def test_input_stage_file_helpers(tmp_path):
    """read_y_component / save_y_frames_to_file / calculate_num_frames (assign1/ex2.py:14-46, common.py:13-19)."""
    import numpy as np
    from basic_video_codec_b200 import input_stage as ist
    from basic_video_codec_b200 import EncoderConfig, InputParameters
    w, h, n = 36, 22, 5
    rng = np.random.default_rng(3)
    ys = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    with open(tmp_path / "clip.yuv", "wb") as fh:
        for y in ys:
            fh.write(y.tobytes())
            fh.write(rng.integers(0, 256, 2 * (w // 2) * (h // 2), dtype=np.uint8).tobytes())
    assert ist.calculate_num_frames(str(tmp_path / "clip.yuv"), w, h) == n
    got = list(ist.read_y_component(str(tmp_path / "clip.yuv"), w, h, n))
    assert all(np.array_equal(a, b) for a, b in zip(got, ys))
    params = InputParameters(str(tmp_path / "clip.y"), w, h, EncoderConfig(4, 2, 2, 3, resolution=(w, h)), frames_to_process=n)
    params.yuv_file = None
    ist.save_y_frames_to_file(params)
    assert open(tmp_path / "clip.y", "rb").read() == ys.tobytes()
    p = ist.pad_frame(ys[0], 8)
    assert p.shape == (24, 40) and np.array_equal(p[:h, :w], ys[0]) and (p[h:, :] == 128).all() and (p[:, w:] == 128).all()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (CPU only): one JSON line with the contract's keys, metric / unit / config of our own arm."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "encoded frames/s" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and "BASELINE configs[3]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 works and prints
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=60, cwd=ROOT, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_division_by_launch_constants_is_exact():
    """The kernels divide by launch constants (tiles per lane, blocks per row, 2r) as multiply-high + add + shift
    (csrc/me_fullsearch.cu div_magic_constants / div_magic, csrc/tq.cu fast_div_constants / fast_div).  Same arithmetic in
    numpy: exact for every divisor the library can meet and every dividend below 2^31."""
    import numpy as np
    rng = np.random.default_rng(5)
    xs = np.concatenate([np.arange(0, 70000, dtype=np.uint64), rng.integers(0, 2**31, 200000, dtype=np.uint64),
                         np.array([2**31 - 1, 2**31 - 2], dtype=np.uint64)])
    for d in list(range(1, 300)) + [510, 1020, 4080, 8160, 65535, 65536, 1 << 20, (1 << 20) + 7]:
        shift = 0
        while (1 << shift) < d:
            shift += 1
        magic = ((1 << 32) * ((1 << shift) - d)) // d + 1
        assert magic < (1 << 32)
        hi = (xs * np.uint64(magic)) >> np.uint64(32)
        s = hi + xs
        assert int(s.max()) < (1 << 32)          # the 32-bit addition in the kernel does not wrap
        assert np.array_equal(s >> np.uint64(shift), xs // np.uint64(d)), d
