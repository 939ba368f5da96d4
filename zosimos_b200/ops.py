"""Direct per-operation entry points of the C-ABI (zos_pixel_chain, zos_compose, ...).  These are
the calls bench.py times and the building blocks the program executor schedules.  Every function
launches CUDA kernels on `ctx`'s stream and returns without synchronising."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from .buffer import Descriptor
from .device import Context, DeviceImage


def texfmt(desc: Descriptor) -> _ffi.ZosTexFmt:
    """The native-vs-staged decision of ImageDescriptor::new (program.rs:781-946), made by the library."""
    lib = _ffi.lib()
    d = desc.to_ffi()
    f = _ffi.ZosTexFmt()
    st = lib.zos_desc_texfmt(C.byref(d), C.byref(f))
    if st != _ffi.OK:
        raise _ffi.ZosError(st, "no texture representation for %r" % (desc,))
    return f


def step(kind: int, m=None, v=None, fmt: Optional[_ffi.ZosTexFmt] = None) -> _ffi.ZosStep:
    s = _ffi.ZosStep()
    s.kind = kind
    if m is not None:
        mm = np.asarray(m, dtype=np.float64).reshape(9)
        for i in range(9):
            s.m[i] = float(np.float32(mm[i]))
    if v is not None:
        for i, x in enumerate(v):
            s.v[i] = float(x)
    if fmt is not None:
        s.fmt = fmt
    return s


def matrix(m) -> _ffi.ZosStep:
    return step(_ffi.STEP_MATRIX, m)


def requant(desc: Descriptor) -> _ffi.ZosStep:
    return step(_ffi.STEP_REQUANT, fmt=texfmt(desc))


def _steps_array(steps: Sequence[_ffi.ZosStep]):
    arr = (_ffi.ZosStep * max(len(steps), 1))()
    for i, s in enumerate(steps):
        arr[i] = s
    return arr


def pixel_chain(ctx: Context, src: DeviceImage, dst: DeviceImage, steps: Sequence[_ffi.ZosStep] = ()):
    a, b = src.ffi(), dst.ffi()
    ctx.check(ctx._lib.zos_pixel_chain(ctx.handle, C.byref(a), C.byref(b), _steps_array(steps), len(steps), dst.batch))


def compose_params(map: int = _ffi.MAP_RECT, sampling: int = _ffi.SAMPLE_NEAREST, blend: int = _ffi.BLEND_OVERWRITE,
                   sel=(0, 0, 0, 0), tgt=(0, 0, 0, 0), inv=None, src_steps=(), dst_steps=(), use_tma: bool = True,
                   dst_origin=(0, 0), src_origin=(0, 0), src_full=(0, 0)) -> _ffi.ZosComposeParams:
    p = _ffi.ZosComposeParams()
    p.map, p.sampling, p.blend, p.use_tma = map, sampling, blend, int(use_tma)
    for i in range(4):
        p.sel[i] = int(sel[i]); p.tgt[i] = int(tgt[i])
    for i in range(2):
        p.dst_origin[i] = int(dst_origin[i]); p.src_origin[i] = int(src_origin[i]); p.src_full[i] = int(src_full[i])
    if inv is not None:
        iv = np.asarray(inv, dtype=np.float32).reshape(9)
        for i in range(9):
            p.inv[i] = float(iv[i])
    p.n_src_steps, p.n_dst_steps = len(src_steps), len(dst_steps)
    for i, s in enumerate(src_steps):
        p.src_steps[i] = s
    for i, s in enumerate(dst_steps):
        p.dst_steps[i] = s
    return p


def compose(ctx: Context, below: Optional[DeviceImage], above: DeviceImage, dst: DeviceImage, params: _ffi.ZosComposeParams):
    b = below.ffi() if below is not None else None
    a, d = above.ffi(), dst.ffi()
    ctx.check(ctx._lib.zos_compose(ctx.handle, C.byref(b) if b is not None else None, C.byref(a), C.byref(d), C.byref(params), dst.batch))


def generate_bilinear(ctx: Context, dst: DeviceImage, params):
    p = (C.c_float * 24)(*[float(x) for x in np.asarray(params, dtype=np.float32).reshape(24)])
    d = dst.ffi()
    ctx.check(ctx._lib.zos_generate_bilinear(ctx.handle, C.byref(d), p, dst.batch))


def generate(ctx: Context, dst: DeviceImage, kind: int, params):
    """zos_generate: kind = _ffi.GEN_*; params as documented in include/zosimos_cuda.h."""
    flat = [float(x) for x in np.asarray(params, dtype=np.float32).reshape(-1)]
    p = (C.c_float * 24)(*(flat + [0.0] * (24 - len(flat))))
    d = dst.ffi()
    ctx.check(ctx._lib.zos_generate(ctx.handle, C.byref(d), int(kind), p, dst.batch))


def box3(ctx: Context, src: DeviceImage, dst: DeviceImage, m):
    mm = (C.c_float * 9)(*[float(x) for x in np.asarray(m, dtype=np.float32).reshape(9)])
    a, d = src.ffi(), dst.ffi()
    ctx.check(ctx._lib.zos_box3(ctx.handle, C.byref(a), C.byref(d), mm, dst.batch))


def palette(ctx: Context, pal: DeviceImage, idx: DeviceImage, dst: DeviceImage, xc, yc):
    x = (C.c_float * 4)(*[float(v) for v in xc]); y = (C.c_float * 4)(*[float(v) for v in yc])
    p, i, d = pal.ffi(), idx.ffi(), dst.ffi()
    ctx.check(ctx._lib.zos_palette(ctx.handle, C.byref(p), C.byref(i), C.byref(d), x, y, dst.batch))
