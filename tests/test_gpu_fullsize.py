"""Parity at BASELINE.json's full sizes.  Where the oracle finishes in seconds the comparison is direct
(C2: 3840x2160 blend / inscribe, C4: one 1080p frame); the large configurations are checked through
size-independent properties (C3 7680x4320: identity and exact-inverse transforms, row-band == whole;
C5 64 MP: round trips and linearity of the byte-exact paths)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.gpu_common import ctx, oracle_desc, oracle_image, zdesc  # noqa: F401

pytestmark = pytest.mark.gpu

import zosimos_b200 as Z  # noqa: E402
from zosimos_b200 import _ffi, ops  # noqa: E402
from zosimos_b200.buffer import Color, SampleParts, Texel, Transfer  # noqa: E402


def test_c2_full_size_blend_and_inscribe_bit_exact(ctx):
    W, H = 3840, 2160
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, (H, W * 4), dtype=np.uint8); b = rng.integers(0, 256, (H, W * 4), dtype=np.uint8)
    a.reshape(H, W, 4)[::7, ::5, 3] = 0; a.reshape(H, W, 4)[3::7, ::3, 3] = 255; b.reshape(H, W, 4)[::5, ::11, 3] = 0
    d = zdesc(W, H, Texel.new_u8(SampleParts.RgbA), Color.SRGB)
    below, above, dst = ctx.upload(d, b), ctx.upload(d, a), ctx.image(d)
    ob, oa = oracle_image(d, b), oracle_image(d, a)
    ops.compose(ctx, below, above, dst, ops.compose_params(blend=_ffi.BLEND_SRC_OVER, sel=(0, 0, W, H), tgt=(0, 0, W, H)))
    assert np.array_equal(dst.download(), O.blend(ob, (0, 0, W, H), oa, 3).data)
    ops.compose(ctx, below, above, dst, ops.compose_params(blend=_ffi.BLEND_OVERWRITE, sel=(0, 0, W, H), tgt=(0, 0, W, H)))
    assert np.array_equal(dst.download(), a)
    # a layer placed inside the canvas (general addressing of the same kernels)
    w2, h2 = 1920, 1080
    a2 = np.ascontiguousarray(a.reshape(H, W, 4)[:h2, :w2]).reshape(h2, w2 * 4)
    d2 = zdesc(w2, h2, Texel.new_u8(SampleParts.RgbA), Color.SRGB)
    above2 = ctx.upload(d2, a2)
    ops.compose(ctx, below, above2, dst, ops.compose_params(blend=_ffi.BLEND_SRC_OVER, sel=(0, 0, w2, h2), tgt=(1000, 500, w2, h2)))
    assert np.array_equal(dst.download(), O.blend(ob, (1000, 500, 1000 + w2, 500 + h2), oracle_image(d2, a2), 3).data)
    for im in (below, above, above2, dst):
        im.free()


def test_c3_full_size_affine_properties(ctx):
    W, H = 7680, 4320
    lin = Color.Rgb(Z.Primaries.Bt709, Transfer.Linear)
    d = zdesc(W, H, Texel.new_f16(), lin)
    rng = np.random.default_rng(3)
    tile = rng.random((135, W * 4), dtype=np.float32).astype(np.float16)
    src = np.tile(tile, (H // 135, 1)).view(np.uint8)
    bel = np.ascontiguousarray(src[::-1])
    above, below, dst = ctx.upload(d, src), ctx.upload(d, bel), ctx.image(d)
    ident = np.eye(3, dtype=np.float32)
    for sampling in (_ffi.SAMPLE_NEAREST, _ffi.SAMPLE_BILINEAR):
        # identity: every destination pixel centre maps onto a source pixel centre: both samplers return the texel
        ops.compose(ctx, below, above, dst, ops.compose_params(map=_ffi.MAP_AFFINE, sampling=sampling, inv=ident, use_tma=True))
        assert np.array_equal(dst.download(), src)
        # integer shift: exact copy of the shifted region, `below` elsewhere
        sx, sy = 1001, 333
        inv = np.array([[1, 0, -sx], [0, 1, -sy], [0, 0, 1]], np.float32)
        ops.compose(ctx, below, above, dst, ops.compose_params(map=_ffi.MAP_AFFINE, sampling=sampling, inv=inv, use_tma=True))
        got = dst.download().reshape(H, W, 8)
        assert np.array_equal(got[sy:, sx:], src.reshape(H, W, 8)[:H - sy, :W - sx])
        assert np.array_equal(got[:sy], bel.reshape(H, W, 8)[:sy]) and np.array_equal(got[:, :sx], bel.reshape(H, W, 8)[:, :sx])
    # rotation by 30 degrees: the dedicated kernel == the general gather kernel, byte for byte, on the whole 33 MP image
    ang = np.deg2rad(30.0)
    m = (O.shift(W / 2, H / 2) @ O.rotate(ang) @ O.shift(-W / 2, -H / 2)).astype(np.float32)
    inv = O.inv3(m.astype(np.float64)).astype(np.float32)
    for sampling in (_ffi.SAMPLE_NEAREST, _ffi.SAMPLE_BILINEAR):
        res = []
        for flags in (0, 1):
            ctx.set_flags(flags)
            ops.compose(ctx, below, above, dst, ops.compose_params(map=_ffi.MAP_AFFINE, sampling=sampling, inv=inv, use_tma=True))
            res.append(dst.download())
        ctx.set_flags(0)
        assert np.array_equal(res[0], res[1])
    # and the oracle on a 256-row window of it (full-image coordinates through the window origin)
    y0, rows = 2048, 256
    band = O.decode(oracle_image(d, bel))[y0:y0 + rows].copy()
    O.paint_affine_window(band, y0, O.decode(oracle_image(d, src)), 0, H, inv, 1)
    enc = O.encode(oracle_desc(zdesc(W, rows, Texel.new_f16(), lin)), band).data
    assert np.array_equal(res[0].reshape(H, W * 8)[y0:y0 + rows], enc)
    for im in (below, above, dst):
        im.free()


def test_c5_full_size_round_trips(ctx):
    W = H = 8192  # 64 MP
    rng = np.random.default_rng(5)
    row = rng.integers(0, 256, (64, W * 4), dtype=np.uint8)
    data = np.tile(row, (H // 64, 1))
    srgb = zdesc(W, H, Texel.new_u8(SampleParts.RgbA), Color.SRGB)
    lin16 = zdesc(W, H, Texel.new_f16(), Color.Rgb(Z.Primaries.Bt709, Transfer.Linear))
    a, mid, back = ctx.upload(srgb, data), ctx.image(lin16), ctx.image(srgb)
    # sRGB8 -> linear RGBA16F -> sRGB8 is the identity (f16 keeps every sRGB8 code apart)
    ops.pixel_chain(ctx, a, mid, [])
    ops.pixel_chain(ctx, mid, back, [])
    assert np.array_equal(back.download(), data)
    # a matrix and its exact inverse permutation: channel swap twice
    P = np.array([[0, 0, 1], [0, 1, 0], [1, 0, 0]], np.float32)
    ops.pixel_chain(ctx, a, back, [ops.matrix(P), ops.matrix(P)])
    assert np.array_equal(back.download(), data)
    # one swap == the BGRA view of the same bytes
    bgra = zdesc(W, H, Texel.new_u8(SampleParts.BgrA), Color.SRGB)
    b2 = ctx.image(bgra)
    ops.pixel_chain(ctx, a, back, [ops.matrix(P)])
    ops.pixel_chain(ctx, a, b2, [])
    assert np.array_equal(back.download(), b2.download())
    for im in (a, mid, back, b2):
        im.free()


def test_c4_full_size_frame_against_oracle(ctx):
    W, H, w, h = 1920, 1080, 1280, 720
    rng = np.random.default_rng(4)
    y = rng.integers(16, 236, (H, W), dtype=np.uint8)
    u = rng.integers(16, 241, (H // 2, W // 2), dtype=np.uint8); v = rng.integers(16, 241, (H // 2, W // 2), dtype=np.uint8)
    yuv = Z.yuv420_descriptor(W, H, Color.Rgb(Z.Primaries.Bt2020, Transfer.Bt709), Z.YuvMatrix.Bt709, False, False, 0)
    od = zdesc(w, h, Texel.new_u8(SampleParts.RgbA), Color.SRGB)
    bgd = rng.integers(0, 256, (h, w * 4), dtype=np.uint8)
    src, bg, dst = ctx.upload(yuv, (y, u, v)), ctx.upload(od, bgd), ctx.image(od)
    M = O.mul3(O.inv3(O.to_xyz("bt709", "D65")), O.to_xyz("bt2020", "D65"))
    p = ops.compose_params(map=_ffi.MAP_SCALE, sampling=_ffi.SAMPLE_BILINEAR, blend=_ffi.BLEND_SRC_OVER, src_steps=[ops.matrix(M)], use_tma=True)
    ops.compose(ctx, bg, src, dst, p)
    got = dst.download()
    tex = O.linear(O.decode_yuv420(y, u, v, W, H, 0.2126, 0.0722, False, False, 0, O.TR_BT709), np.array(M, np.float32).reshape(3, 3))
    exp = O.encode(oracle_desc(od), O.resize_pass(tex, w, h, 1)).data   # the frame is opaque: source-over == the frame
    dd = np.abs(got.astype(int) - exp.astype(int))
    assert dd.max() <= 1 and np.mean(dd == 0) > 0.99   # pow on the SFU vs libm
    res2 = []
    for flags in (0, 1):                                # specialised kernel == general kernels, byte for byte
        ctx.set_flags(flags)
        ops.compose(ctx, bg, src, dst, p)
        res2.append(dst.download())
    ctx.set_flags(0)
    assert np.array_equal(res2[0], res2[1])
    for im in (src, bg, dst):
        im.free()
