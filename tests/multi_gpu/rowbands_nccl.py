"""Row-band sharding of ONE large image over N GPUs + NCCL gather of the bands, run as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu/rowbands_nccl.py

Every rank resamples its band of output rows from the source rows that band needs (full-image coordinates through
the window origins), the bands are all-gathered over NVLink, and rank 0 checks that the assembled image equals
the single-GPU whole-image result byte for byte.  (pytest does not collect this file; the partition arithmetic
itself is covered on CPU with gloo in tests/test_shard.py.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import zosimos_b200 as Z  # noqa: E402
from oracle import oracle as O  # noqa: E402  (matrix helpers only)
from zosimos_b200 import _ffi, ops  # noqa: E402
from zosimos_b200.buffer import ByteLayout, Color, Descriptor, Texel, Transfer  # noqa: E402
from zosimos_b200.shard import band_source_rows, gather_outputs, row_bands  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Z.Context(local)
    W, H = 4096, 4096
    lin = Color.Rgb(Z.Primaries.Bt709, Transfer.Linear)

    def desc(w, h):
        return Descriptor(ByteLayout(w, h, w * 8, 8), lin, Texel.new_f16())

    rng = np.random.default_rng(5)  # same data on every rank (the source is replicated)
    src = rng.random((H, W * 4), dtype=np.float32).astype(np.float16).view(np.uint8)
    bel = np.ascontiguousarray(src[::-1])
    ang = np.deg2rad(17.0)
    m = (O.shift(W / 2, H / 2) @ O.rotate(ang) @ O.scale(1.1, 0.9) @ O.shift(-W / 2, -H / 2)).astype(np.float32)
    inv = O.inv3(m.astype(np.float64)).astype(np.float32)
    bands = row_bands(H, world, 32)
    assert all(b[1] - b[0] == bands[0][1] - bands[0][0] for b in bands), "equal bands for all_gather"
    y0, y1 = bands[rank]
    s0, s1 = band_source_rows(inv.reshape(9), (y0, y1), W, H)
    above = ctx.upload(desc(W, s1 - s0), src[s0:s1])
    below = ctx.upload(desc(W, y1 - y0), bel[y0:y1])
    dst = ctx.image(desc(W, y1 - y0))
    p = ops.compose_params(map=_ffi.MAP_AFFINE, sampling=_ffi.SAMPLE_BILINEAR, inv=inv, use_tma=True,
                           dst_origin=(0, y0), src_origin=(0, s0), src_full=(W, H))
    ops.compose(ctx, below, above, dst, p)
    band = torch.from_numpy(dst.download().copy()).cuda()
    parts = gather_outputs(band)
    if rank == 0:
        whole_a, whole_b, whole_d = ctx.upload(desc(W, H), src), ctx.upload(desc(W, H), bel), ctx.image(desc(W, H))
        ops.compose(ctx, whole_b, whole_a, whole_d, ops.compose_params(map=_ffi.MAP_AFFINE, sampling=_ffi.SAMPLE_BILINEAR, inv=inv, use_tma=True))
        exp = whole_d.download()
        got = torch.cat(parts, 0).cpu().numpy()
        ok = np.array_equal(got, exp)
        print("row-band sharding over %d GPU(s) + NCCL all_gather: %s (%d x %d RGBA16F, bands of %d rows)" % (world, "byte-identical to the single-GPU image" if ok else "MISMATCH", W, H, y1 - y0), flush=True)
        if not ok:
            sys.exit(1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
