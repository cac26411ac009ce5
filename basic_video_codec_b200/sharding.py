"""GOP sharding across ranks: an I frame clears the reference window
(reference encoder/encoder.py:174-178), so GOPs are independent units.  GOP g goes to rank
g mod world; every rank encodes its GOPs on its own GPU; there is no collective on the data path.
The per-GOP container fragments are gathered on rank 0 and concatenated in GOP order, which is
byte-identical to the serial stream (tests/test_gpu_parity.py::test_gop_streams_concatenate).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def gop_ranges(nframes: int, i_period: int) -> List[Tuple[int, int]]:
    """[(first_frame, n_frames)] of every GOP of a clip that starts on an I frame."""
    return [(f0, min(i_period, nframes - f0)) for f0 in range(0, nframes, i_period)]


def assign_gops(ngop: int, world: int) -> List[List[int]]:
    """Round-robin GOP -> rank map (rank r gets g with g % world == r)."""
    return [list(range(r, ngop, world)) for r in range(world)]


def scaling_ceiling(ngop: int, world: int) -> float:
    """Best possible strong-scaling speed-up: the busiest rank encodes ceil(ngop/world) GOPs."""
    return ngop / float(-(-ngop // world))


def _default_encode(frames: np.ndarray, ec, device: int) -> bytes:
    from .clip import encode_clip
    return encode_clip(frames, ec, device=device)[0]


def encode_clip_distributed(frames: Optional[np.ndarray], ec, *, rank: int = 0, world: int = 1, device: int = 0,
                            encode_fn: Optional[Callable[[np.ndarray, object, int], bytes]] = None,
                            load_gop: Optional[Callable[[int, int], np.ndarray]] = None,
                            nframes: Optional[int] = None) -> Optional[bytes]:
    """Encode a clip with GOPs sharded over `world` ranks (torch.distributed must be initialised when
    world > 1; any backend -- only gather_object is used).  Every rank passes either the whole `frames`
    array or a `load_gop(first, n)` callback so that it only touches its own GOPs.  Returns the
    container bytes on rank 0, None elsewhere."""
    encode_fn = encode_fn or _default_encode
    n = int(nframes if nframes is not None else frames.shape[0])
    ranges = gop_ranges(n, ec.I_Period)
    mine = assign_gops(len(ranges), world)[rank]
    parts = []
    if mine:
        # all GOPs of this rank in one call so the GPU encodes them in lock-step lanes
        chunks = [(load_gop(*ranges[g]) if load_gop else frames[ranges[g][0]: ranges[g][0] + ranges[g][1]]) for g in mine]
        full = [c for c, g in zip(chunks, mine) if ranges[g][1] == ec.I_Period]
        if len(full) == len(chunks):
            data = encode_fn(np.concatenate(chunks, axis=0), ec, device)
            parts = split_container_by_gop(data, [ranges[g][1] for g in mine])
        else:  # a short last GOP: encode it on its own
            parts = [encode_fn(c, ec, device) for c in chunks]
    payload = list(zip(mine, parts))
    if world == 1:
        gathered = [payload]
    else:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(payload, gathered, dst=0)
        if rank != 0:
            return None
    ordered = sorted((g, p) for lst in gathered for g, p in lst)
    return b"".join(p for _, p in ordered)


def split_container_by_gop(data: bytes, frames_per_gop: Sequence[int]) -> List[bytes]:
    """Cut a container (reference encoder/encoder.py:104-121: mode | len16 | pred | len24 | coef per frame)
    into per-GOP fragments."""
    out, o = [], 0
    for nf in frames_per_gop:
        start = o
        for _ in range(nf):
            pl = int.from_bytes(data[o + 1:o + 3], "big")
            o += 3 + pl
            cl = int.from_bytes(data[o:o + 3], "big")
            o += 3 + cl
        out.append(data[start:o])
    if o != len(data):
        raise ValueError("container length does not match the GOP structure")
    return out
