"""Two contexts on two devices in ONE process (the library's per-device state: constant tables, dynamic shared
memory limits, NVRTC modules, tensor maps).  Run on a box with >= 2 GPUs: python tests/multi_gpu/two_devices_one_process.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import zosimos_b200 as Z  # noqa: E402
from oracle import oracle as O  # noqa: E402
from zosimos_b200 import _ffi, ops  # noqa: E402
from zosimos_b200.buffer import ByteLayout, Color, Descriptor, SampleParts, Texel, Transfer  # noqa: E402


def run(ctx):
    out = {}
    rng = np.random.default_rng(1)
    W, H = 640, 360
    srgb = Descriptor(ByteLayout(W, H, W * 4, 4), Color.SRGB, Texel.new_u8(SampleParts.RgbA))
    a = rng.integers(0, 256, (H, W * 4), dtype=np.uint8); b = rng.integers(0, 256, (H, W * 4), dtype=np.uint8)
    A, B, D = ctx.upload(srgb, a), ctx.upload(srgb, b), ctx.image(srgb)
    ops.compose(ctx, B, A, D, ops.compose_params(blend=_ffi.BLEND_SRC_OVER, sel=(0, 0, W, H), tgt=(0, 0, W, H)))
    out["blend"] = D.download()
    T = O.to_xyz("bt709", "D65")
    lch = Descriptor(ByteLayout(W, H, W * 4, 4), Color.Oklab, Texel(Z.Block.Pixel, Z.SampleBits.UInt8x4, SampleParts.LchA))
    ops.pixel_chain(ctx, A, D, [ops.step(_ffi.STEP_OKLAB_ENC, T), ops.requant(lch), ops.step(_ffi.STEP_OKLAB_DEC, O.inv3(T))])
    out["lab"] = D.download()
    f16 = Descriptor(ByteLayout(W, H, W * 8, 8), Color.Rgb(Z.Primaries.Bt709, Transfer.Linear), Texel.new_f16())
    src = rng.random((H, W * 4), dtype=np.float32).astype(np.float16).view(np.uint8)
    S, Bl, Dd = ctx.upload(f16, src), ctx.upload(f16, np.ascontiguousarray(src[::-1])), ctx.image(f16)
    m = (O.shift(W / 2, H / 2) @ O.rotate(0.4) @ O.shift(-W / 2, -H / 2)).astype(np.float32)
    inv = O.inv3(m.astype(np.float64)).astype(np.float32)
    ops.compose(ctx, Bl, S, Dd, ops.compose_params(map=_ffi.MAP_AFFINE, sampling=_ffi.SAMPLE_BILINEAR, inv=inv, use_tma=True))
    out["affine"] = Dd.download()
    y = rng.integers(16, 236, (H, W), dtype=np.uint8)
    u = rng.integers(16, 241, (H // 2, W // 2), dtype=np.uint8); v = rng.integers(16, 241, (H // 2, W // 2), dtype=np.uint8)
    yuv = Z.yuv420_descriptor(W, H, Color.Rgb(Z.Primaries.Bt2020, Transfer.Bt709), Z.YuvMatrix.Bt709, False, False, 0)
    small = Descriptor(ByteLayout(426, 240, 426 * 4, 4), Color.SRGB, Texel.new_u8(SampleParts.RgbA))
    Y, Ds = ctx.upload(yuv, (y, u, v)), ctx.image(small)
    ops.compose(ctx, None, Y, Ds, ops.compose_params(map=_ffi.MAP_SCALE, sampling=_ffi.SAMPLE_BILINEAR, use_tma=True))
    out["frame"] = Ds.download()
    return out


def main():
    c0, c1 = Z.Context(0), Z.Context(1)
    r1 = run(c1)   # device 1 FIRST: state configured for device 0 only would show here
    r0 = run(c0)
    r1b = run(c1)
    ok = all(np.array_equal(r0[k], r1[k]) and np.array_equal(r1[k], r1b[k]) for k in r0)
    print("two devices in one process:", "identical results on both" if ok else "MISMATCH", sorted(r0))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
