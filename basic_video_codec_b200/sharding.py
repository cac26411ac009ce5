"""GOP sharding across ranks: an I frame clears the reference window
(reference encoder/encoder.py:174-178), so GOPs are independent units.  Every rank encodes a contiguous run of
GOPs on its own GPU; there is no collective on the data path -- not even for the bookkeeping: what the ranks exchange
is one integer each (the length of their container fragment), through mailboxes in a host buffer shared by the ranks
of the node (a /dev/shm mapping, page-locked).  Every rank then copies its fragment from device memory straight to
its offset in that buffer, so the serial stream
-- byte-identical to a single-GPU encode, the reference's encoded.bin layout (encoder.py:104-121) -- is written
exactly once and rank 0 returns a view of it.
"""
from __future__ import annotations

import mmap
import os
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def gop_ranges(nframes: int, i_period: int) -> List[Tuple[int, int]]:
    """[(first_frame, n_frames)] of every GOP of a clip that starts on an I frame."""
    return [(f0, min(i_period, nframes - f0)) for f0 in range(0, nframes, i_period)]


def assign_gops(ngop: int, world: int) -> List[List[int]]:
    """GOP -> rank map: contiguous runs, the first ngop % world ranks hold one GOP more.  Contiguous, so a rank's
    container fragment is one contiguous slice of the serial stream."""
    base, extra = divmod(ngop, world)
    out, g = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(list(range(g, g + n)))
        g += n
    return out


def scaling_ceiling(ngop: int, world: int) -> float:
    """Best possible strong-scaling speed-up: the busiest rank encodes ceil(ngop/world) GOPs."""
    return ngop / float(-(-ngop // world))


_HDR = 4096      # bytes in front of the stream: per-rank mailboxes (see ShardedEncoder._exchange)


class _SharedBuffer:
    """A host buffer all ranks of the node map: a file in /dev/shm created by rank 0 (name broadcast once).  The first _HDR
    bytes are the ranks' mailboxes, the stream follows."""

    def __init__(self, nbytes: int, rank: int, world: int):
        self.nbytes = int(nbytes) + _HDR
        self.path = None
        self.owner = rank == 0
        if world == 1:
            self.mm = mmap.mmap(-1, self.nbytes)
        else:
            import torch.distributed as dist
            name = [None]
            if rank == 0:
                self.path = f"/dev/shm/bvc_shard_{os.getpid()}_{int.from_bytes(os.urandom(4), 'little'):08x}"
                fd = os.open(self.path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                os.ftruncate(fd, self.nbytes)
                name[0] = self.path
            dist.broadcast_object_list(name, src=0)
            if rank != 0:
                self.path = name[0]
                fd = os.open(self.path, os.O_RDWR)
            self.mm = mmap.mmap(fd, self.nbytes)
            os.close(fd)
            dist.barrier()           # everyone has it mapped: the name can go
            if rank == 0:
                os.unlink(self.path)
        whole = np.frombuffer(self.mm, dtype=np.uint8)
        self.mail = whole[:_HDR].view(np.int64).reshape(-1, 4)   # [rank] = (sequence, fragment bytes, done sequence, -)
        if rank == 0:
            self.mail[:] = 0
        self.array = whole[_HDR:]
        self.registered = False
        if world > 1:
            import torch.distributed as dist
            dist.barrier()           # mailboxes zeroed before anyone posts

    def close(self):
        if self.registered:
            from ._lib import host_unregister
            host_unregister(self.array)
            self.registered = False
        self.array = None
        self.mail = None
        try:
            self.mm.close()
        except BufferError:      # a view handed out by encode() is still alive: the mapping goes with it
            pass


class ShardedEncoder:
    """One clip, GOPs sharded over `world` ranks (torch.distributed initialised when world > 1, any backend).
    Build once per (clip geometry, EncoderConfig); encode() may be called repeatedly.

    encode_fn(frames, ec, device) -> bytes replaces the GPU encoder (the CPU tests pass the oracle)."""

    def __init__(self, ec, width: int, height: int, nframes: int, *, rank: int = 0, world: int = 1, device: int = 0,
                 encode_fn: Optional[Callable] = None, capacity: Optional[int] = None):
        if getattr(ec, "RCflag", 0) in (2, 3):
            raise NotImplementedError("RCflag 2/3 couple consecutive GOPs (prev_frame.rc_qp_per_row): replicas only")
        self.ec, self.W, self.H, self.nframes = ec, int(width), int(height), int(nframes)
        self.rank, self.world, self.device, self.encode_fn = rank, world, device, encode_fn
        self.ranges = gop_ranges(self.nframes, ec.I_Period)
        self.mine = assign_gops(len(self.ranges), world)[rank]
        self.first = self.ranges[self.mine[0]][0] if self.mine else 0
        self.count = sum(self.ranges[g][1] for g in self.mine)
        self.capacity = int(capacity or (self.nframes * self.W * self.H // 2 + (1 << 20)))
        self.buf = _SharedBuffer(self.capacity, rank, world)
        self.ctx = None
        self._seq = 0
        if encode_fn is None:
            from ._lib import Context, host_register
            if self.mine:
                self.ctx = Context(self.W, self.H, ec.block_size, ec.search_range, ec.quantization_factor, ec.nRefFrames,
                                   ec.fastME, ec.fracMeEnabled, ec.I_Period, device=device, max_lanes=len(self.mine))
                from .clip import configure_rate_control
                configure_rate_control(self.ctx, ec)     # RCflag 1 couples rows of one frame only: GOPs stay independent
            host_register(self.buf.array)
            self.buf.registered = True

    # ---- the one exchange: fragment lengths, through the mailboxes in the shared buffer ----
    # All ranks run on one node (they share the host buffer), so the exchange is two 8-byte stores per rank and a spin on
    # the other ranks' sequence numbers: a few microseconds, no collective, no GPU work.  (x86 / aarch64 store order
    # within one writer: the length is written before the sequence number that announces it.)
    def _post_length(self, n: int) -> List[int]:
        self._seq += 1
        if self.world == 1:
            return [n]
        mail = self.buf.mail
        mail[self.rank, 1] = n
        mail[self.rank, 0] = self._seq
        deadline = None
        while True:
            if bool((mail[:self.world, 0] >= self._seq).all()):
                return [int(x) for x in mail[:self.world, 1]]
            deadline = self._spin(deadline)

    def _post_done(self):
        if self.world == 1:
            return
        mail = self.buf.mail
        mail[self.rank, 2] = self._seq
        if self.rank != 0:
            return
        deadline = None
        while not bool((mail[:self.world, 2] >= self._seq).all()):   # rank 0 hands out the stream: every fragment is in place
            deadline = self._spin(deadline)

    @staticmethod
    def _spin(deadline):
        import time
        now = time.monotonic()
        if deadline is None:
            return now + 600.0
        if now > deadline:
            raise TimeoutError("a rank of the sharded encode did not post its fragment within 10 minutes")
        return deadline

    def my_frames(self, frames: Optional[np.ndarray], load_gop: Optional[Callable[[int, int], np.ndarray]] = None):
        """The frames of this rank's GOPs: a view of `frames`, or load_gop(first, n) per GOP (only its own GOPs touched)."""
        if not self.mine:
            return None
        if load_gop is not None:
            chunks = [load_gop(*self.ranges[g]) for g in self.mine]
            return chunks[0] if len(chunks) == 1 else np.concatenate(chunks, axis=0)
        return frames[self.first:self.first + self.count]

    def encode(self, frames: Optional[np.ndarray] = None, load_gop: Optional[Callable[[int, int], np.ndarray]] = None,
               resident: bool = False):
        """Encode the clip.  Returns a uint8 view of the whole container on rank 0, None elsewhere (the view is valid until
        the next encode() of any rank).  resident=True encodes what upload() put in HBM (throughput measurements without the
        input transfer)."""
        data = None
        if not self.mine:
            n = 0
        elif self.encode_fn is not None:
            data = self.encode_fn(self.my_frames(frames, load_gop), self.ec, self.device)
            n = len(data)
        elif resident:
            n = self.ctx.encode_clip_device(None, self.count)
        else:
            n = self.ctx.encode_clip_device(self.my_frames(frames, load_gop))
        sizes = self._post_length(n)
        total, off = sum(sizes), sum(sizes[:self.rank])
        if total > self.capacity:
            raise MemoryError(f"container of {total} bytes exceeds the shared buffer ({self.capacity})")
        if n:
            if data is not None:
                self.buf.array[off:off + n] = np.frombuffer(data, dtype=np.uint8)
            else:
                self.ctx.container_download(self.buf.array, off, 0, n)
        self._post_done()
        return self.buf.array[:total] if self.rank == 0 else None

    def upload(self, frames: Optional[np.ndarray] = None, load_gop=None):
        if self.ctx is not None:
            self.ctx.clip_upload(self.my_frames(frames, load_gop))

    def close(self):
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None
        if self.buf is not None:
            self.buf.close()
            self.buf = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def encode_clip_distributed(frames: Optional[np.ndarray], ec, *, rank: int = 0, world: int = 1, device: int = 0,
                            encode_fn: Optional[Callable[[np.ndarray, object, int], bytes]] = None,
                            load_gop: Optional[Callable[[int, int], np.ndarray]] = None,
                            nframes: Optional[int] = None, shape: Optional[Tuple[int, int]] = None) -> Optional[bytes]:
    """Encode a clip with GOPs sharded over `world` ranks.  Every rank passes either the whole `frames` array or a
    `load_gop(first, n)` callback (+ nframes and shape=(H, W)) so that it only touches its own GOPs.  Returns the
    container bytes on rank 0, None elsewhere."""
    n = int(nframes if nframes is not None else frames.shape[0])
    H, W = shape if shape is not None else frames.shape[1:]
    with ShardedEncoder(ec, W, H, n, rank=rank, world=world, device=device, encode_fn=encode_fn) as enc:
        out = enc.encode(frames, load_gop)
        return bytes(out) if out is not None else None


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
