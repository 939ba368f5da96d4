"""Device context and images over the C-ABI: the device half of the reference's Pool
(/root/reference/lib/zosimos/src/pool.rs:39-41 `Gpu`, :122-156 `ImageData::GpuBuffer`) and the
host<->device row copies of its executor (lib/zosimos/src/run.rs:3282-3309, 2265-2270).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _ffi
from .buffer import Block, Descriptor


class Context:
    """One CUDA device + stream (zos_ctx).  Not thread safe, like `&mut Execution`."""

    def __init__(self, device: int = 0):
        self._lib = _ffi.lib()
        h = C.c_void_p()
        st = self._lib.zos_ctx_create(int(device), C.byref(h))
        if st != _ffi.OK:
            raise _ffi.ZosError(st, (self._lib.zos_last_error(None) or b"").decode())
        self.handle = h
        self.device = int(device)

    def check(self, st: int):
        if st != _ffi.OK:
            raise _ffi.ZosError(st, (self._lib.zos_last_error(self.handle) or b"").decode())

    @property
    def stream(self) -> int:
        return int(self._lib.zos_ctx_stream(self.handle) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.zos_ctx_launch_count(self.handle))

    def set_flags(self, flags: int):
        """ZOS_CTX_NO_FAST_PATHS = 1: route every launch through the generic kernels (parity checks)."""
        self.check(self._lib.zos_ctx_set_flags(self.handle, int(flags)))

    def sync(self):
        self.check(self._lib.zos_sync(self.handle))

    def poll(self) -> bool:
        """zos_poll: True once everything enqueued on this context has finished; never blocks."""
        done = C.c_int32(0)
        self.check(self._lib.zos_poll(self.handle, C.byref(done)))
        return bool(done.value)

    def arena_stats(self) -> dict:
        """zos_ctx_arena_stats: cudaMalloc calls so far, allocations served from parked blocks, bytes held / in use / parked."""
        st = _ffi.ZosArenaStats()
        self.check(self._lib.zos_ctx_arena_stats(self.handle, C.byref(st)))
        return {k: int(getattr(st, k)) for k, _ in _ffi.ZosArenaStats._fields_}

    def arena_trim(self):
        """Pool::clear_cache for this device (pool.rs:450-455): parked blocks go back to the driver."""
        self.check(self._lib.zos_ctx_arena_trim(self.handle))

    def close(self):
        if self.handle:
            self._lib.zos_ctx_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- memory
    def alloc(self, nbytes: int) -> "DeviceBuffer":
        return DeviceBuffer(self, nbytes)

    def pinned(self, nbytes: int) -> "PinnedArray":
        return PinnedArray(self, nbytes)

    def image(self, desc: Descriptor, batch: int = 1) -> "DeviceImage":
        return DeviceImage(self, desc, batch)

    def upload(self, desc: Descriptor, data, batch: int = 1) -> "DeviceImage":
        img = DeviceImage(self, desc, batch)
        img.upload(data)
        return img


class DeviceBuffer:
    def __init__(self, ctx: Context, nbytes: int):
        self.ctx = ctx
        h = C.c_void_p()
        ctx.check(ctx._lib.zos_buf_alloc(ctx.handle, int(nbytes), C.byref(h)))
        self.handle = h
        self.nbytes = int(nbytes)
        self.ptr = int(ctx._lib.zos_buf_ptr(h) or 0)

    def free(self):
        if self.handle and self.ctx.handle:
            self.ctx._lib.zos_buf_free(self.ctx.handle, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """Page-locked host staging memory (the map_write/map_read buffers of encoder.rs:574-616)."""

    def __init__(self, ctx: Context, nbytes: int):
        self.ctx = ctx
        p = C.c_void_p()
        ctx.check(ctx._lib.zos_host_alloc(ctx.handle, int(nbytes), C.byref(p)))
        self.ptr = p
        self.nbytes = int(nbytes)
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(p.value))

    def free(self):
        if self.ptr:
            self.array = None
            self.ctx._lib.zos_host_free(self.ctx.handle, self.ptr)
            self.ptr = None


class DeviceImage:
    """A (batch of) image(s) in a pitched device buffer; rows padded to 256 bytes like
    Descriptor::to_aligned (buffer.rs:121-134).  Planar 4:2:0 frames keep Y, U, V (or Y, UV) in one
    allocation."""

    def __init__(self, ctx: Context, desc: Descriptor, batch: int = 1):
        self.ctx, self.desc, self.batch = ctx, desc, int(batch)
        w, h = desc.size()
        lib = ctx._lib
        self.planar = desc.texel.block != Block.Pixel
        if not self.planar:
            self.pitch = desc.to_aligned().row_stride
            self.frame_bytes = self.pitch * h
            self.cpitch = self.cframe = 0
        else:
            nv12 = desc.texel.block == Block.Yuv420Nv12
            self.cw, self.ch = (w + 1) // 2, (h + 1) // 2
            self.pitch = int(lib.zos_aligned_row_stride(w, 1))
            self.cpitch = int(lib.zos_aligned_row_stride(self.cw * (2 if nv12 else 1), 1))
            self.y_bytes = self.pitch * h
            self.c_bytes = self.cpitch * self.ch
            self.frame_bytes = self.y_bytes + self.c_bytes * (1 if nv12 else 2)
        self.buf = DeviceBuffer(ctx, self.frame_bytes * self.batch)

    def ffi(self) -> _ffi.ZosImage:
        im = _ffi.ZosImage()
        im.desc = self.desc.to_ffi(self.pitch)
        im.data = self.buf.ptr
        im.batch_stride = self.frame_bytes if self.batch > 1 else 0
        if self.planar:
            im.plane1 = self.buf.ptr + self.y_bytes
            im.plane2 = self.buf.ptr + self.y_bytes + self.c_bytes if self.desc.texel.block == Block.Yuv420Planar else None
            im.chroma_stride = self.cpitch
            im.chroma_batch_stride = self.frame_bytes if self.batch > 1 else 0
        return im

    # ---- host <-> device.  Host arrays: (batch, h, row_bytes) uint8 tight rows; planar: dict/tuple of planes
    def _rows(self, data, rows, row_bytes):
        a = np.ascontiguousarray(data, dtype=np.uint8).reshape(self.batch, rows, row_bytes)
        return a

    def upload(self, data):
        lib, ctx = self.ctx._lib, self.ctx
        w, h = self.desc.size()
        if not self.planar:
            rb = w * self.desc.layout.texel_stride
            a = self._rows(data, h, rb)
            for f in range(self.batch):
                ctx.check(lib.zos_buf_upload(ctx.handle, self.buf.handle, f * self.frame_bytes, self.pitch,
                                             a[f].ctypes.data_as(C.c_void_p), rb, rb, h))
        else:
            y, u, v = data  # NV12: u = interleaved plane, v = None
            nv12 = self.desc.texel.block == Block.Yuv420Nv12
            y = self._rows(y, h, w)
            crb = self.cw * (2 if nv12 else 1)
            u = self._rows(u, self.ch, crb)
            vv = None if nv12 else self._rows(v, self.ch, crb)
            for f in range(self.batch):
                base = f * self.frame_bytes
                ctx.check(lib.zos_buf_upload(ctx.handle, self.buf.handle, base, self.pitch, y[f].ctypes.data_as(C.c_void_p), w, w, h))
                ctx.check(lib.zos_buf_upload(ctx.handle, self.buf.handle, base + self.y_bytes, self.cpitch,
                                             u[f].ctypes.data_as(C.c_void_p), crb, crb, self.ch))
                if vv is not None:
                    ctx.check(lib.zos_buf_upload(ctx.handle, self.buf.handle, base + self.y_bytes + self.c_bytes, self.cpitch,
                                                 vv[f].ctypes.data_as(C.c_void_p), crb, crb, self.ch))
        ctx.sync()  # the numpy source is pageable: the copy has completed, keep the contract simple

    def download(self) -> np.ndarray:
        lib, ctx = self.ctx._lib, self.ctx
        w, h = self.desc.size()
        if self.planar:  # (y, u, v) planes of frame 0..batch-1; NV12: (y, interleaved uv, None)
            nv12 = self.desc.texel.block == Block.Yuv420Nv12
            crb = self.cw * (2 if nv12 else 1)
            y = np.empty((self.batch, h, w), np.uint8)
            u = np.empty((self.batch, self.ch, crb), np.uint8)
            v = None if nv12 else np.empty((self.batch, self.ch, crb), np.uint8)
            for f in range(self.batch):
                base = f * self.frame_bytes
                ctx.check(lib.zos_buf_download(ctx.handle, self.buf.handle, base, self.pitch, y[f].ctypes.data_as(C.c_void_p), w, w, h))
                ctx.check(lib.zos_buf_download(ctx.handle, self.buf.handle, base + self.y_bytes, self.cpitch,
                                               u[f].ctypes.data_as(C.c_void_p), crb, crb, self.ch))
                if v is not None:
                    ctx.check(lib.zos_buf_download(ctx.handle, self.buf.handle, base + self.y_bytes + self.c_bytes, self.cpitch,
                                                   v[f].ctypes.data_as(C.c_void_p), crb, crb, self.ch))
            ctx.sync()
            if self.batch == 1:
                return y[0], u[0], (None if v is None else v[0])
            return y, u, v
        rb = w * self.desc.layout.texel_stride
        out = np.empty((self.batch, h, rb), np.uint8)
        for f in range(self.batch):
            ctx.check(lib.zos_buf_download(ctx.handle, self.buf.handle, f * self.frame_bytes, self.pitch,
                                           out[f].ctypes.data_as(C.c_void_p), rb, rb, h))
        ctx.sync()
        return out[0] if self.batch == 1 else out

    def host_frame_bytes(self) -> int:
        """Bytes of one frame in the tight host layout (plane 0 rows, then the chroma planes)."""
        w, h = self.desc.size()
        if not self.planar:
            return w * h * self.desc.layout.texel_stride
        return w * h + 2 * self.cw * self.ch

    def upload_from(self, host_ptr: int, sync: bool = True):
        """Every frame from tight host memory at `host_ptr` (zos_image_upload).  Asynchronous on the context's stream when
        the memory is pinned and `sync` is False."""
        im, ctx, fb = self.ffi(), self.ctx, self.host_frame_bytes()
        for f in range(self.batch):
            ctx.check(ctx._lib.zos_image_upload(ctx.handle, C.byref(im), f, C.c_void_p(host_ptr + f * fb)))
        if sync:
            ctx.sync()

    def download_into(self, host_ptr: int, sync: bool = True):
        im, ctx, fb = self.ffi(), self.ctx, self.host_frame_bytes()
        for f in range(self.batch):
            ctx.check(ctx._lib.zos_image_download(ctx.handle, C.byref(im), f, C.c_void_p(host_ptr + f * fb)))
        if sync:
            ctx.sync()

    def free(self):
        self.buf.free()
