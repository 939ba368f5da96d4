"""The multi-GPU entry points of the C-ABI (include/zosimos_cuda.h) as far as ONE device can exercise them:
two contexts on device 0 stand for two devices (zos_multi_launch, zos_multi_sync, zos_gather_peer take the same code path except
for the peer copy itself), and a one-rank NCCL communicator runs zos_comm_* / zos_gather_nccl end to end (dlopen of NCCL, init,
all-gather and send/recv forms).  The real thing -- 2 and 8 GPUs, byte-identity with the single-GPU image -- is
tests/multi_gpu/gather_c_abi.py (profiles/r02_multi_gpu.md)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import zosimos_b200 as Z  # noqa: E402
from zosimos_b200 import _ffi, ops, shard  # noqa: E402
from zosimos_b200.buffer import ByteLayout, Color, Descriptor, SampleParts, Texel  # noqa: E402


def _desc(w, h):
    return Descriptor(ByteLayout(w, h, w * 4, 4), Color.SRGB, Texel.new_u8(SampleParts.RgbA))


def _blend_program(ctx, below, above, dst):
    """input, input, compose(source-over), output as a zos_program with everything bound."""
    w, h = dst.desc.size()
    arr = (_ffi.ZosOp * 4)()
    for i, im in enumerate((below, above)):
        arr[i].kind = _ffi.OP_INPUT; arr[i].src[0] = arr[i].src[1] = -1; arr[i].dst = i; arr[i].reg = i; arr[i].desc = im.ffi().desc
    o = arr[2]
    o.kind = _ffi.OP_COMPOSE; o.src[0] = 0; o.src[1] = 1; o.dst = 2; o.reg = 2; o.desc = dst.ffi().desc
    o.compose = ops.compose_params(blend=_ffi.BLEND_SRC_OVER, sel=(0, 0, w, h), tgt=(0, 0, w, h))
    o = arr[3]
    o.kind = _ffi.OP_OUTPUT; o.src[0] = 2; o.src[1] = -1; o.dst = 2; o.reg = 3; o.desc = dst.ffi().desc
    p = C.c_void_p()
    ctx.check(ctx._lib.zos_program_create(ctx.handle, arr, 4, _ffi.FUSE_EXACT, 1, C.byref(p)))
    for reg, im in ((0, below), (1, above), (2, dst)):
        f = im.ffi()
        ctx.check(ctx._lib.zos_program_bind(p, reg, C.byref(f)))
    return p


def test_multi_launch_and_gather_peer_between_two_contexts():
    W, H = 512, 96  # two bands of 48 rows, one per context
    rng = np.random.default_rng(11)
    a = rng.integers(0, 256, (H, W * 4), dtype=np.uint8); b = rng.integers(0, 256, (H, W * 4), dtype=np.uint8)
    ctxs = [Z.Context(0), Z.Context(0)]
    try:
        whole = ctxs[0].image(_desc(W, H))
        ops.compose(ctxs[0], ctxs[0].upload(_desc(W, H), b), ctxs[0].upload(_desc(W, H), a), whole,
                    ops.compose_params(blend=_ffi.BLEND_SRC_OVER, sel=(0, 0, W, H), tgt=(0, 0, W, H)))
        expect = whole.download()
        bands = shard.row_bands(H, 2, 16)
        jobs, progs = [], []
        for c, (y0, y1) in zip(ctxs, bands):
            d = _desc(W, y1 - y0)
            below, above, dst = c.upload(d, b[y0:y1]), c.upload(d, a[y0:y1]), c.image(d)
            jobs.append((below, above, dst))
            progs.append(_blend_program(c, below, above, dst))
        full = ctxs[0].alloc(H * W * 4)
        for graph in (False, True, True):  # eager, then the captured graph twice
            shard.multi_launch(progs, graph=graph)
            shard.gather_peer(ctxs[0], full, [y0 * W * 4 for y0, _ in bands], [(c, j[2].buf, 0, (y1 - y0) * W * 4) for c, j, (y0, y1) in zip(ctxs, jobs, bands)])
            shard.multi_sync(ctxs)
            got = np.empty((H, W * 4), np.uint8)
            ctxs[0].check(ctxs[0]._lib.zos_buf_download(ctxs[0].handle, full.handle, 0, W * 4, C.c_void_p(got.ctypes.data), W * 4, W * 4, H))
            ctxs[0].sync()
            assert np.array_equal(got, expect)
        # out-of-range shards are refused, nothing is copied
        with pytest.raises(_ffi.ZosError):
            shard.gather_peer(ctxs[0], full, [H * W * 4 - 16], [(ctxs[1], jobs[1][2].buf, 0, 4096)])
        for c, p in zip(ctxs, progs):
            c._lib.zos_program_destroy(p)
    finally:
        for c in ctxs:
            c.close()


def test_gather_nccl_with_one_rank():
    lib = _ffi.lib()
    if lib.zos_comm_nccl_version() == 0:
        pytest.skip("NCCL cannot be loaded in this process")
    with Z.Context(0) as ctx:
        comm = shard.Comm(ctx, 0, 1, shard.Comm.unique_id())
        n = 1 << 20
        src, dst = ctx.alloc(n), ctx.alloc(2 * n)
        data = np.random.default_rng(3).integers(0, 256, n, dtype=np.uint8)
        ctx.check(lib.zos_buf_upload(ctx.handle, src.handle, 0, n, C.c_void_p(data.ctypes.data), n, n, 1))
        for root, off in ((-1, 0), (0, n)):  # all-gather form, then the send / recv form (only the rank's own shard: a device copy)
            comm.gather(src, 0, dst, [off], [n], root)
            got = np.empty(n, np.uint8)
            ctx.check(lib.zos_buf_download(ctx.handle, dst.handle, off, n, C.c_void_p(got.ctypes.data), n, n, 1))
            ctx.sync()
            assert np.array_equal(got, data)
        with pytest.raises(_ffi.ZosError):
            comm.gather(src, 0, dst, [2 * n - 8], [n], -1)  # does not fit
        comm.close()
