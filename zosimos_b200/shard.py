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
