"""The C oracle's texel unpack / pack against an independent numpy restatement of the bit layouts of
lib/std/src/stage.frag:533-641 (`demux_uint` / `mux_uint`; codes lib/zosimos/src/shaders/stage.rs:98-119), as tabulated in
SURVEY.md appendix A.3: fields are listed from the least significant bit, a field of n bits decodes to k / (2^n - 1) in f32
(then rests in the f16 working texture) and packs the f16 texture value as uint(clamp(c, 0, 1) * (2^n - 1)) --
truncation, not rounding.
Integer work: every comparison here is exact."""
import numpy as np
import pytest

from oracle import oracle as O

# bits code -> (word dtype, [(slot in (x, y, z, w), bit width), ...] from the least significant bit)
LAYOUTS = {
    "UInt332": (O.B_UINT332, np.uint8, [(0, 2), (1, 3), (2, 3)]),
    "UInt233": (O.B_UINT233, np.uint8, [(0, 3), (1, 3), (2, 2)]),
    "UInt4x4": (O.B_UINT4X4, np.uint16, [(0, 4), (1, 4), (2, 4), (3, 4)]),
    "UInt565": (O.B_UINT565, np.uint16, [(0, 5), (1, 6), (2, 5)]),
    "UInt8x4": (O.B_UINT8X4, np.uint32, [(0, 8), (1, 8), (2, 8), (3, 8)]),
    "UInt2101010": (O.B_UINT2101010, np.uint32, [(0, 2), (1, 10), (2, 10), (3, 10)]),
    "UInt1010102": (O.B_UINT1010102, np.uint32, [(0, 10), (1, 10), (2, 10), (3, 2)]),  # RGB10A2: R low, A in the top 2 bits
}


def words_for(dtype, rng):
    if dtype == np.uint8:
        return np.arange(256, dtype=np.uint8)
    if dtype == np.uint16:
        return np.arange(65536, dtype=np.uint16)
    edge = np.array([0, 0xFFFFFFFF, 0x3FF, 0xFFC00, 0x3FF00000, 0xC0000000, 0x80000000, 1], np.uint32)
    return np.concatenate([edge, rng.integers(0, 2 ** 32, 65536 - len(edge), dtype=np.uint32)])


def parts_of(fields):
    return O.P_RGBA if len(fields) == 4 else O.P_RGB


@pytest.mark.parametrize("name", sorted(LAYOUTS))
def test_unpack_fields(name):
    bits, dtype, fields = LAYOUTS[name]
    words = words_for(dtype, np.random.default_rng(7))
    desc = O.Desc(len(words), 1, O.Texel(bits, parts_of(fields)), O.SCALARS_LINEAR)
    got = O.decode(O.Image(desc, words.view(np.uint8).reshape(1, -1)))[0]
    exp = np.ones((len(words), 4), np.float32)  # absent fields read 1 (w of the 3-field formats)
    shift = 0
    for slot, n in fields:
        k = (words.astype(np.uint64) >> np.uint64(shift)) & np.uint64((1 << n) - 1)
        exp[:, slot] = (k.astype(np.float32) / np.float32((1 << n) - 1)).astype(np.float16).astype(np.float32)
        shift += n
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("name", sorted(LAYOUTS))
def test_pack_truncates(name):
    bits, dtype, fields = LAYOUTS[name]
    rng = np.random.default_rng(11)
    n_px = 40000
    tex = rng.uniform(-0.25, 1.25, (1, n_px, 4)).astype(np.float32)
    # field boundaries k / max as the f16 working texture holds them, and their f16 neighbours: where rounding and
    # truncation differ the most
    for slot, n in fields:
        m = (1 << n) - 1
        k = rng.integers(0, m + 1, 3000)
        b = (k / m).astype(np.float16)
        tex[0, :3000, slot] = b
        tex[0, 3000:6000, slot] = np.nextafter(b, np.float16(-1))
        tex[0, 6000:9000, slot] = np.nextafter(b, np.float16(2))
    desc = O.Desc(n_px, 1, O.Texel(bits, parts_of(fields)), O.SCALARS_LINEAR)
    got = O.encode(desc, tex).data.view(dtype).reshape(-1)
    exp = np.zeros(n_px, np.uint64)
    shift = 0
    for slot, n in fields:
        c = np.clip(tex[0, :, slot].astype(np.float16).astype(np.float32), np.float32(0), np.float32(1))  # the draw wrote f16
        exp |= (c * np.float32((1 << n) - 1)).astype(np.uint64) << np.uint64(shift)  # f32 product, then truncation
        shift += n
    assert np.array_equal(got.astype(np.uint64), exp)


def test_two_field_quirks():
    """UInt8x2 holds (x, w); UInt16x2 DECODES its second field into y and reads w = 1 (stage.frag:557-558) although the
    encoder writes w there (stage.frag:592,619): LumaA16 therefore loses its alpha on the way in.  Bgra is swizzled by the
    encoder only (stage.frag:676-677 vs 725-726)."""
    w8 = np.array([(200 << 8) | 100], np.uint16)
    got = O.decode(O.Image(O.Desc(1, 1, O.Texel(O.B_UINT8X2, O.P_LUMAA), O.SCALARS_LINEAR), w8.view(np.uint8).reshape(1, -1)))[0, 0]
    f16 = lambda v: np.float32(np.float16(np.float32(v)))
    assert list(got) == [f16(np.float32(100) / np.float32(255))] * 3 + [f16(np.float32(200) / np.float32(255))]
    w16 = np.array([(2000 << 16) | 1000], np.uint32)
    got = O.decode(O.Image(O.Desc(1, 1, O.Texel(O.B_UINT16X2, O.P_LUMAA), O.SCALARS_LINEAR), w16.view(np.uint8).reshape(1, -1)))[0, 0]
    assert list(got) == [f16(np.float32(1000) / np.float32(65535))] * 3 + [np.float32(1)]
    tex = np.array([[[0.25, 0.5, 0.75, 1.0]]], np.float32)
    enc = O.encode(O.Desc(1, 1, O.Texel(O.B_UINT16X2, O.P_LUMAA), O.SCALARS_LINEAR), tex).data.view(np.uint32)[0, 0]
    assert (int(enc) & 0xFFFF, int(enc) >> 16) == (int(np.float32(0.25) * np.float32(65535)), 65535)
    px = np.array([0x04030201], np.uint32).view(np.uint8).reshape(1, -1)
    rgba = O.decode(O.Image(O.Desc(1, 1, O.Texel(O.B_UINT8X4, O.P_RGBA), O.SCALARS_LINEAR), px))
    bgra = O.decode(O.Image(O.Desc(1, 1, O.Texel(O.B_UINT8X4, O.P_BGRA), O.SCALARS_LINEAR), px))
    assert np.array_equal(rgba, bgra)
    enc = O.encode(O.Desc(1, 1, O.Texel(O.B_UINT8X4, O.P_BGRA), O.SCALARS_LINEAR), np.array([[[0.0, 0.5, 1.0, 1.0]]], np.float32)).data[0]
    assert list(enc) == [255, 127, 0, 255]
