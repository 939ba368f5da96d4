"""Partitioning of the hot path across GPUs: the path shards into independent units
with NO exchange step -- frames of a batch, or row bands of one very large image -- so every rank
(one process per GPU) runs the same kernels on its own slice; the only optional collective is a
gather of the outputs (NCCL all-gather through torch.distributed; gloo in the CPU tests).

The reference has nothing comparable: its Pool can hold several devices but always picks the first
(lib/zosimos/src/pool.rs:227-230; "If multi-device then this should become a set", run.rs:416-420).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np


def bind_to_gpu_numa(device: int) -> Optional[List[int]]:
    """One process per GPU: pin this process to the CPUs that share the GPU's PCIe root (its NUMA node)
    BEFORE allocating pinned host memory, so that staging buffers are local to the GPU.  Without it the
    end-to-end (host buffer) throughput of 8 ranks collapses onto one socket's memory and interconnect.
    Returns the CPU list, or None when the topology cannot be read (then nothing is changed)."""
    import os
    import subprocess
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(device), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bdf.startswith("00000000:"):
            bdf = "0000:" + bdf[9:]
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            spec = f.read().strip()
        cpus: List[int] = []
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.extend(range(int(lo), int(hi) + 1))
            elif part:
                cpus.append(int(part))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None


def frame_shard(n_frames: int, rank: int, world: int) -> range:
    """Contiguous, balanced partition: rank r owns frames [r*n/world, (r+1)*n/world)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return range(rank * n_frames // world, (rank + 1) * n_frames // world)


def row_bands(height: int, world: int, align: int = 1) -> List[Tuple[int, int]]:
    """Splits `height` output rows into `world` bands [y0, y1) whose starts are multiples of `align`
    (32 = the gather kernel's tile height keeps tiles whole).  Bands may be empty for tiny images."""
    units = (height + align - 1) // align
    out = []
    for r in range(world):
        y0 = min(r * units // world * align, height)
        y1 = min((r + 1) * units // world * align, height)
        out.append((y0, y1))
    return out


def band_source_rows(inv: Sequence[float], band: Tuple[int, int], dst_width: int, src_height: int, margin: int = 2) -> Tuple[int, int]:
    """Source rows [sy0, sy1) an output band needs under the inverse affine map `inv` (row-major 3x3,
    destination pixel centre -> source coordinates): the map is linear, so the extremes are at the
    band's corners; `margin` covers the bilinear footprint and rounding.  Never empty for a non-empty source."""
    y0, y1 = band
    if y1 <= y0:
        return (0, 0)
    m = np.asarray(inv, dtype=np.float64).reshape(3, 3)
    ys = []
    for cx in (0.5, dst_width - 0.5):
        for cy in (y0 + 0.5, y1 - 0.5):
            ys.append(m[1, 0] * cx + m[1, 1] * cy + m[1, 2])
    lo = min(max(int(math.floor(min(ys))) - margin, 0), src_height)  # a band that maps entirely below the source gets an empty window at its end
    hi = min(int(math.floor(max(ys))) + margin + 1, src_height)
    if hi <= lo and src_height > 0:  # the band does not touch the source: one (unread) row, so that callers never upload an empty image
        lo = min(lo, src_height - 1)
        hi = lo + 1
    return (lo, max(hi, lo))


class Comm:
    """zos_comm: an NCCL communicator bound to a context (one process per GPU).  `id128` is made by rank 0
    (`Comm.unique_id()`) and handed to the other ranks by any host channel; `from_torch_distributed` uses the process
    group that is already up (gloo or nccl) for that hand-over only -- the gather itself is zos_gather_nccl on the
    context's stream."""

    def __init__(self, ctx, rank: int, world: int, id128: bytes):
        import ctypes as C
        from . import _ffi
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(id128))
        ctx.check(ctx._lib.zos_comm_create(ctx.handle, buf, self.rank, self.world, C.byref(h)))
        self.handle = h

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from . import _ffi
        lib = _ffi.lib()
        buf = (C.c_uint8 * 128)()
        st = lib.zos_comm_unique_id(buf)
        if st != _ffi.OK:
            raise _ffi.ZosError(st, (lib.zos_last_error(None) or b"").decode())
        return bytes(buf)

    @staticmethod
    def from_torch_distributed(ctx, group=None) -> "Comm":
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return Comm(ctx, rank, world, box[0])

    def gather(self, send, send_off: int, recv, recv_offsets: Sequence[int], shard_bytes: Sequence[int], root: int = -1):
        """Every rank contributes shard_bytes[rank] bytes of `send` (a DeviceBuffer) from send_off; rank `root` (every rank
        when root < 0) receives shard r at recv + recv_offsets[r].  Asynchronous on the context's stream."""
        import ctypes as C
        n = self.world
        offs = (C.c_uint64 * n)(*[int(v) for v in recv_offsets]) if recv_offsets is not None else None
        sizes = (C.c_uint64 * n)(*[int(v) for v in shard_bytes])
        self.ctx.check(self.ctx._lib.zos_gather_nccl(self.handle, send.handle if send is not None else None, int(send_off),
                                                     recv.handle if recv is not None else None, offs, sizes, int(root)))

    def close(self):
        if self.handle:
            self.ctx._lib.zos_comm_destroy(self.handle)
            self.handle = None


def gather_peer(dst_ctx, dst, dst_offsets: Sequence[int], shards):
    """zos_gather_peer for contexts of ONE process: shards = [(ctx, DeviceBuffer, offset, bytes), ...] copied over NVLink
    into `dst` (a DeviceBuffer of dst_ctx) at dst_offsets; ordered after each source context's queued work."""
    import ctypes as C
    n = len(shards)
    ctxs = (C.c_void_p * n)(*[s[0].handle for s in shards])
    bufs = (C.c_void_p * n)(*[s[1].handle for s in shards])
    soff = (C.c_uint64 * n)(*[int(s[2]) for s in shards])
    size = (C.c_uint64 * n)(*[int(s[3]) for s in shards])
    doff = (C.c_uint64 * n)(*[int(v) for v in dst_offsets])
    dst_ctx.check(dst_ctx._lib.zos_gather_peer(dst_ctx.handle, dst.handle, doff, ctxs, bufs, soff, size, n))


def multi_launch(programs, graph: bool = True):
    """zos_multi_launch: start the (already bound) zos_program handles of several contexts from one host thread."""
    import ctypes as C
    from . import _ffi
    lib = _ffi.lib()
    n = len(programs)
    arr = (C.c_void_p * n)(*programs)
    st = lib.zos_multi_launch(arr, n, _ffi.RUN_GRAPH if graph else _ffi.RUN_EAGER)
    if st != _ffi.OK:
        raise _ffi.ZosError(st, "zos_multi_launch")


def multi_sync(ctxs):
    import ctypes as C
    from . import _ffi
    n = len(ctxs)
    arr = (C.c_void_p * n)(*[c.handle for c in ctxs])
    st = _ffi.lib().zos_multi_sync(arr, n)
    if st != _ffi.OK:
        raise _ffi.ZosError(st, "; ".join((c._lib.zos_last_error(c.handle) or b"").decode() for c in ctxs))


def gather_outputs(local, group=None):
    """All-gather of per-rank outputs (a torch tensor; equal shapes) -> list of tensors, rank order.
    Outside of any timed region: the data path itself has no collective."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return [local]
    world = dist.get_world_size(group)
    outs = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(outs, local.contiguous(), group=group)
    return outs
