"""The C oracle's transfer functions (f32) against an independent float64 numpy restatement of the reference's formulas
(lib/std/src/stage.frag:280-425 and the dispatch at :749-809), including what the reference does differently from the
standards it cites: Bt601 is the BT.709 pair SWAPPED (:308-315), the SMPTE 2084 exponent m2 lacks the factor 128
(:367) and its decode uses the plain EOTF while its encode goes through the OOTF (:400-405, :768, :800), Bt2100Pq and
Bt2100Hlg fall through as the identity.  Float texels (Float32x4) expose the functions without quantisation.
Tolerance: 1e-5 relative (+ 1e-6 absolute), the float bound of BASELINE.json's north star."""
import numpy as np
import pytest

from oracle import oracle as O


F = lambda c: float(np.float32(c))  # the shader compares f32 values with f32 constants


def oe709(v):
    return np.where(v >= F(0.018), 1.099 * np.power(np.maximum(v, 0.0), 0.45) - 0.099, 4.5 * v)


def eo709(v):
    thr = F(1.099 * 0.018 ** 0.45 - 0.099)
    return np.where(v >= thr, np.power((np.maximum(v, thr) + 0.099) / 1.099, 1 / 0.45), v / 4.5)


def oe240(v):
    return np.where(v < F(0.0228), 4.0 * v, 1.1115 * np.power(np.maximum(v, 0.0), 0.45) - 0.1115)


def eo240(v):
    with np.errstate(invalid="ignore"):  # [0.0913, 0.1115) is NaN in the reference: pow of a negative base (stage.frag:327-332)
        return np.where(v < F(0.0913), v / 4.0, np.power((v - 0.1115) / 1.1115, 1 / 0.45))


def oe_srgb(v):
    a = np.abs(v)
    return np.sign(v) * np.where(a <= F(0.0031308), a * 12.92, 1.055 * np.power(a, 1 / 2.4) - 0.055)


def eo_srgb(v):
    a = np.abs(v)
    return np.sign(v) * np.where(a <= F(0.04045), a / 12.92, np.power((a + 0.055) / 1.055, 2.4))


M1, M2, C1, C2, C3 = 2610 / 16384, 2523 / 4096, 3424 / 4096, 2413 / 128, 2392 / 128  # m2 as the shader has it


def eo_pq(v):
    n = np.power(v, 1 / M2)
    return np.power(np.maximum(n - C1, 0) / (C2 - C3 * n), 1 / M1)


def oe_pq(v):  # eo_inv(scene -> display)
    y = np.power(np.power(oe709(59.5208 * v), 2.4) / 100.0, M1)
    return np.power((C1 + C2 * y) / (C3 * y + 1.0), M2)


ident = lambda v: v
PAIRS = {  # transfer code -> (decode = parts_untransfer, encode = parts_transfer)
    O.TR_BT709: (eo709, oe709), O.TR_BT470M: (lambda v: np.power(v, 2.2), lambda v: np.power(v, 1 / 2.2)),
    O.TR_BT601: (oe709, eo709), O.TR_SMPTE240: (eo240, oe240), O.TR_LINEAR: (ident, ident), O.TR_SRGB: (eo_srgb, oe_srgb),
    O.TR_BT2020_10: (eo709, oe709), O.TR_BT2020_12: (eo709, oe709), O.TR_SMPTE2084: (eo_pq, oe_pq),
    O.TR_BT2100PQ: (ident, ident), O.TR_BT2100HLG: (ident, ident),
}


def samples(signed):
    v = np.concatenate([np.linspace(0, 1, 4097), [0.018, 0.0228, 0.0913, 0.0031308, 0.04045, 0.081, 1e-4, 1e-6],
                        np.nextafter(np.float32([0.018, 0.0228, 0.0913, 0.0031308, 0.04045]), np.float32(0)).astype(np.float64)])
    if signed:
        v = np.concatenate([v, -v[1:]])
    return v.astype(np.float32)


def close(got, exp, rel=1e-5):
    got = got.astype(np.float64)
    return np.all((np.abs(got - exp) <= rel * np.abs(exp) + 1e-6) | (np.isnan(got) & np.isnan(exp)))


@pytest.mark.parametrize("tr", sorted(PAIRS))
def test_transfer_pair(tr):
    v = samples(signed=tr == O.TR_SRGB)  # sRGB is odd-symmetric (Y'CbCr excursions); the others are used on [0, 1]
    tex = np.stack([v, v[::-1], v, np.linspace(0, 1, len(v), dtype=np.float32)], -1)[None]
    desc = O.Desc(len(v), 1, O.Texel(O.B_FLOAT32X4, O.P_RGBA), O.Color("rgb", tr))
    dec_f, enc_f = PAIRS[tr]
    dec = O.decode(O.Image(desc, np.ascontiguousarray(tex).view(np.uint8).reshape(1, -1)))[0]
    enc = O.encode(desc, tex).data.view(np.float32).reshape(-1, 4)
    for ch in range(3):
        x = tex[0, :, ch].astype(np.float64)
        # the PQ EOTF cancels in both its numerator and denominator near 1: 1e-4 is what an f32 evaluation keeps there
        assert close(dec[:, ch], dec_f(x), 1e-4 if tr == O.TR_SMPTE2084 else 1e-5), (tr, "decode", ch)
        assert close(enc[:, ch], enc_f(x)), (tr, "encode", ch)
    assert np.array_equal(dec[:, 3], tex[0, :, 3]) and np.array_equal(enc[:, 3], tex[0, :, 3])  # alpha is never transferred


def test_lab_lch():  # stage.frag:407-418: C = |ab|, h = atan2(b, a) / 360 deg + 1/2, and back
    rng = np.random.default_rng(5)
    lab = np.concatenate([rng.uniform(0, 1, (4000, 1)), rng.uniform(-0.4, 0.4, (4000, 2)), rng.uniform(0, 1, (4000, 1))], 1).astype(np.float32)
    desc = O.Desc(4000, 1, O.Texel(O.B_FLOAT32X4, O.P_LCHA), O.OKLAB)
    lch = O.encode(desc, lab[None]).data.view(np.float32).reshape(-1, 4)
    L, a, b = (lab[:, i].astype(np.float64) for i in range(3))
    assert close(lch[:, 0], L) and close(lch[:, 1], np.hypot(a, b))
    dh = np.abs(lch[:, 2].astype(np.float64) - (np.arctan2(b, a) / (2 * np.pi) + 0.5))
    assert np.all(np.minimum(dh, 1 - dh) <= 2e-6)
    back = O.decode(O.Image(desc, np.ascontiguousarray(lch).view(np.uint8).reshape(1, -1)))[0]
    assert np.all(np.abs(back.astype(np.float64) - lab) <= 2e-6)


def test_native_srgb8_texture_path():
    """Native Rgba8UnormSrgb (program.rs:794-838): the texture unit decodes sRGB8 exactly (the correctly rounded f32 of the
    piecewise formula) and encodes with round-to-nearest; alpha is linear unorm8.  All 256 codes, and the rounding boundary
    between neighbouring codes located to one f32 step."""
    k = np.arange(256, dtype=np.uint8)
    px = np.stack([k, k, k, k], -1).reshape(1, -1)
    desc = O.srgb_rgba8(256, 1)
    tex = O.decode(O.Image(desc, px))[0]
    exp = eo_srgb(k.astype(np.float64) / 255.0)
    assert np.array_equal(tex[:, 0], exp.astype(np.float32)) and np.array_equal(tex[:, 3], (k.astype(np.float64) / 255).astype(np.float32))
    assert np.array_equal(O.encode(desc, tex[None]).data.reshape(-1, 4), px.reshape(-1, 4))  # round trip of every code
    # the boundary between codes c and c + 1 is where oe_srgb(v) * 255 crosses c + 1/2
    for c in (0, 1, 10, 11, 127, 200, 254):
        lo, hi = eo_srgb(c / 255.0), eo_srgb((c + 1) / 255.0)
        for _ in range(60):
            mid = (lo + hi) / 2
            lo, hi = (mid, hi) if oe_srgb(np.float64(mid)) * 255 < c + 0.5 else (lo, mid)
        b = np.float32(lo)
        below, above = np.nextafter(b, np.float32(-1), dtype=np.float32), np.nextafter(np.nextafter(b, np.float32(2), dtype=np.float32), np.float32(2), dtype=np.float32)
        t = np.array([[[below, above, 0, 1]]], np.float32)
        enc = O.encode(O.srgb_rgba8(1, 1), t).data[0]
        assert (enc[0], enc[1]) == (c, c + 1), (c, enc[:2])
