"""The multi-GPU entry points of the C-ABI on real devices:

  one process per GPU   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu/gather_c_abi.py
                        row bands of one 8192 x 8192 RGBA16F affine resample, gathered with zos_gather_nccl (all ranks, then
                        root only, then ragged bands through ncclSend / ncclRecv); rank 0 compares with the whole-image run.
  one process, N GPUs   python tests/multi_gpu/gather_c_abi.py --peer N
                        the same bands as N programs started by zos_multi_launch, gathered with zos_gather_peer.

Prints the gather time (CUDA events, max over ranks) next to the kernel time.  pytest does not collect this file; the partition
arithmetic is covered on CPU with gloo in tests/test_shard.py."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import zosimos_b200 as Z  # noqa: E402
from zosimos_b200 import _ffi, ops, shard  # noqa: E402
from zosimos_b200.buffer import ByteLayout, Color, Descriptor, Texel, Transfer  # noqa: E402
from zosimos_b200.command import Affine, AffineSample  # noqa: E402

W = H = 8192
LIN = Color.Rgb(Z.Primaries.Bt709, Transfer.Linear)


def desc(w, h):
    return Descriptor(ByteLayout(w, h, w * 8, 8), LIN, Texel.new_f16())


def source():
    rng = np.random.default_rng(5)  # the same data on every rank (the source is replicated)
    tile = rng.random((256, W * 4), dtype=np.float32).astype(np.float16)
    src = np.tile(tile, (H // 256, 1)).view(np.uint8)
    return src, np.ascontiguousarray(src[::-1])


def inverse():
    a = Affine.new(AffineSample.Nearest).shift(-W / 2, -H / 2).rotate(float(np.deg2rad(17.0))).scale(1.1, 0.9).shift(W / 2, H / 2)
    m = np.asarray(a.transformation, dtype=np.float64).reshape(3, 3)
    return np.linalg.inv(m).astype(np.float32)


def band_job(ctx, src, bel, inv, band):
    y0, y1 = band
    s0, s1 = shard.band_source_rows(inv.reshape(9), band, W, H)
    above, below, dst = ctx.upload(desc(W, s1 - s0), src[s0:s1]), ctx.upload(desc(W, y1 - y0), bel[y0:y1]), ctx.image(desc(W, y1 - y0))
    p = ops.compose_params(map=_ffi.MAP_AFFINE, sampling=_ffi.SAMPLE_BILINEAR, inv=inv, use_tma=True, dst_origin=(0, y0), src_origin=(0, s0), src_full=(W, H))
    return above, below, dst, p


def whole(ctx, src, bel, inv):
    a, b, d = ctx.upload(desc(W, H), src), ctx.upload(desc(W, H), bel), ctx.image(desc(W, H))
    ops.compose(ctx, b, a, d, ops.compose_params(map=_ffi.MAP_AFFINE, sampling=_ffi.SAMPLE_BILINEAR, inv=inv, use_tma=True))
    out = d.download()
    for im in (a, b, d):
        im.free()
    return out


def main_nccl():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Z.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    comm = shard.Comm.from_torch_distributed(ctx)
    src, bel = source()
    inv = inverse()
    ok = True
    report = []
    for label, bands in (("equal bands", shard.row_bands(H, world, 32)),
                         ("ragged bands", [(0, 0)] * 0 + _ragged(world))):
        empty = bands[rank][1] <= bands[rank][0]  # a rank without rows still takes part in the gather (0 bytes)
        above, below, dst, p = band_job(ctx, src, bel, inv, bands[rank] if not empty else (0, 32))
        sizes = [(b[1] - b[0]) * W * 8 for b in bands]
        offs = [b[0] * W * 8 for b in bands]
        full = ctx.alloc(H * W * 8)
        for root in (-1, 0):
            times = []
            for it in range(4):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ctx.sync(); dist.barrier()
                e[0].record(stream)
                if not empty:
                    ops.compose(ctx, below, above, dst, p)
                e[1].record(stream)
                comm.gather(dst.buf, 0, full if (root < 0 or rank == root) else None, offs if (root < 0 or rank == root) else None, sizes, root)
                e[2].record(stream)
                ctx.sync()
                t = torch.tensor([e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times.append(t.tolist())
            k_ms, g_ms = min(t[0] for t in times), min(t[1] for t in times)
            if root < 0 or rank == root:
                got = np.empty(H * W * 8, np.uint8)
                ctx.check(ctx._lib.zos_buf_download(ctx.handle, full.handle, 0, W * 8, C.c_void_p(got.ctypes.data), W * 8, W * 8, H))
                ctx.sync()
                if rank == 0:
                    exp = whole(ctx, src, bel, inv)
                    same = np.array_equal(got.reshape(H, W * 8), exp)
                    ok = ok and same
                    report.append("%s, gather to %s: %s; kernel %.3f ms, gather %.3f ms (%.1f GB/s into each receiver)" % (
                        label, "every rank" if root < 0 else "rank 0", "byte-identical to the single-GPU image" if same else "MISMATCH", k_ms, g_ms,
                        (H * W * 8 - sizes[rank]) / g_ms / 1e6))
        for im in (above, below, dst):
            im.free()
        full.free()
    if rank == 0:
        print("zos_gather_nccl over %d GPU(s), NCCL %d, %d x %d RGBA16F:" % (world, ctx._lib.zos_comm_nccl_version(), W, H))
        for r in report:
            print("  " + r)
        sys.stdout.flush()
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


def _ragged(world):
    """Unequal bands (multiples of 32 rows; the first rank gets the most, an empty band when there are many ranks)."""
    if world == 1:
        return [(0, H)]
    weights = np.arange(world, 0, -1, dtype=np.float64)
    weights[-1] = 0.0 if world > 2 else 1.0
    edges = np.concatenate([[0], np.cumsum(weights) / weights.sum() * (H // 32)]).round().astype(int) * 32
    return [(int(edges[i]), int(edges[i + 1])) for i in range(world)]


def main_peer(n):
    ctxs = [Z.Context(i) for i in range(n)]
    src, bel = source()
    inv = inverse()
    bands = shard.row_bands(H, n, 32)
    jobs = [band_job(c, src, bel, inv, b) for c, b in zip(ctxs, bands)]
    # the bands as programs: one compose op each, started together by zos_multi_launch
    progs = []
    for c, (above, below, dst, p) in zip(ctxs, jobs):
        opsarr = (_ffi.ZosOp * 4)()
        for i, (kind, im) in enumerate(((_ffi.OP_INPUT, below), (_ffi.OP_INPUT, above))):
            opsarr[i].kind = kind; opsarr[i].src[0] = opsarr[i].src[1] = -1; opsarr[i].dst = i; opsarr[i].reg = i; opsarr[i].desc = im.ffi().desc
        o = opsarr[2]
        o.kind = _ffi.OP_COMPOSE; o.src[0] = 0; o.src[1] = 1; o.dst = 2; o.reg = 2; o.desc = dst.ffi().desc; o.compose = p
        o = opsarr[3]
        o.kind = _ffi.OP_OUTPUT; o.src[0] = 2; o.src[1] = -1; o.dst = 2; o.reg = 3; o.desc = dst.ffi().desc
        h = C.c_void_p()
        c.check(c._lib.zos_program_create(c.handle, opsarr, 4, _ffi.FUSE_EXACT, 1, C.byref(h)))
        for reg, im in ((0, below), (1, above), (2, dst)):
            f = im.ffi()
            c.check(c._lib.zos_program_bind(h, reg, C.byref(f)))
        progs.append(h)
    full = ctxs[0].alloc(H * W * 8)
    sizes = [(b[1] - b[0]) * W * 8 for b in bands]
    best = None
    for it in range(4):
        shard.multi_sync(ctxs)
        t0 = time.perf_counter()
        shard.multi_launch(progs, graph=False)
        shard.gather_peer(ctxs[0], full, [b[0] * W * 8 for b in bands], [(c, j[2].buf, 0, s) for c, j, s in zip(ctxs, jobs, sizes)])
        shard.multi_sync(ctxs)
        dt = (time.perf_counter() - t0) * 1e3
        best = dt if best is None else min(best, dt)
    got = np.empty(H * W * 8, np.uint8)
    ctxs[0].check(ctxs[0]._lib.zos_buf_download(ctxs[0].handle, full.handle, 0, W * 8, C.c_void_p(got.ctypes.data), W * 8, W * 8, H))
    ctxs[0].sync()
    for c, h in zip(ctxs, progs):
        c._lib.zos_program_destroy(h)
    exp = whole(ctxs[0], src, bel, inv)
    same = np.array_equal(got.reshape(H, W * 8), exp)
    print("zos_multi_launch + zos_gather_peer, %d GPU(s) in one process, %d x %d RGBA16F: %s; launch + gather + sync %.3f ms (host clock)" % (
        n, W, H, "byte-identical to the single-GPU image" if same else "MISMATCH", best), flush=True)
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--peer":
        main_peer(int(sys.argv[2]))
    else:
        main_nccl()
