"""Property tests of the oracle for the semantics that have NO reference implementation (SURVEY.md 8c /
DESIGN.md section 3: Porter-Duff blending, bilinear sampling, planar YUV 4:2:0, float texels).  The
reference's goldens cannot pin these, so the definitions are checked against the algebra they are meant
to implement: identities, the classic Porter-Duff table, associativity of `over`, round trips.  CPU only.
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle as O

SET = settings(max_examples=25, deadline=None)


def rgba(rng, h, w, alpha=None):
    t = rng.random((h, w, 4), dtype=np.float32)
    if alpha is not None:
        t[..., 3] = alpha
    return t


def premul(t):
    return np.concatenate([t[..., :3].astype(np.float64) * t[..., 3:4], t[..., 3:4].astype(np.float64)], -1)


# ---------------------------------------------------------------- Porter-Duff (command.rs:1510-1519 is UNIMPLEMENTED)
@SET
@given(seed=st.integers(0, 2**31 - 1), mode=st.integers(0, 11))
def test_porter_duff_matches_the_premultiplied_table(seed, mode):
    """co = Fa * cs + Fb * cd on premultiplied colours with the classic (Fa, Fb) pairs (Porter & Duff 1984,
    table 1); the oracle works on straight alpha with one reciprocal, so compare after premultiplying."""
    rng = np.random.default_rng(seed)
    s, d = rgba(rng, 5, 7), rgba(rng, 5, 7)
    if seed % 3 == 0:
        s[0, :, 3] = 0.0; d[1, :, 3] = 0.0; s[2, :, 3] = 1.0; d[3, :, 3] = 1.0
    out = O.blend_pass(d.copy(), s, 0, 0, mode)
    a_s, a_d = s[..., 3:4].astype(np.float64), d[..., 3:4].astype(np.float64)
    fa, fb = {0: (0 * a_s, 0 * a_s), 1: (1 + 0 * a_s, 0 * a_s), 2: (0 * a_s, 1 + 0 * a_s), 3: (1 + 0 * a_s, 1 - a_s), 4: (1 - a_d, 1 + 0 * a_s),
              5: (a_d, 0 * a_s), 6: (0 * a_s, a_s), 7: (1 - a_d, 0 * a_s), 8: (0 * a_s, 1 - a_s), 9: (a_d, 1 - a_s), 10: (1 - a_d, a_s),
              11: (1 - a_d, 1 - a_s)}[mode]
    exp = fa * premul(s) + fb * premul(d)
    assert np.allclose(premul(out), exp, atol=2e-6)
    assert np.all(out[..., :3][out[..., 3] == 0] == 0)  # alpha 0 -> colour 0, not NaN


@SET
@given(seed=st.integers(0, 2**31 - 1))
def test_over_identities(seed):
    rng = np.random.default_rng(seed)
    d = rgba(rng, 4, 6)
    clear = rgba(rng, 4, 6, alpha=0.0)
    opaque = rgba(rng, 4, 6, alpha=1.0)
    assert np.array_equal(O.blend_pass(d.copy(), clear, 0, 0, 3)[..., 3], d[..., 3])       # transparent source: alpha kept exactly
    assert np.allclose(O.blend_pass(d.copy(), clear, 0, 0, 3), d, rtol=3e-7, atol=1e-7)    # ... colour within one rounding of c * a / a
    assert np.array_equal(O.blend_pass(d.copy(), opaque, 0, 0, 3), opaque)                  # opaque source replaces, exactly
    assert np.array_equal(O.blend_pass(d.copy(), opaque, 0, 0, 1), opaque)                  # `src`
    assert np.allclose(O.blend_pass(d.copy(), opaque, 0, 0, 2), d, rtol=3e-7, atol=1e-7)   # `dst`
    assert np.all(O.blend_pass(d.copy(), opaque, 0, 0, 0) == 0)                             # `clear`


@SET
@given(seed=st.integers(0, 2**31 - 1))
def test_over_is_associative(seed):
    rng = np.random.default_rng(seed)
    a, b, c = rgba(rng, 3, 5), rgba(rng, 3, 5), rgba(rng, 3, 5)
    left = O.blend_pass(O.blend_pass(c.copy(), b, 0, 0, 3), a, 0, 0, 3)        # a over (b over c)
    ab = O.blend_pass(b.copy(), a, 0, 0, 3)
    right = O.blend_pass(c.copy(), ab, 0, 0, 3)                                # (a over b) over c
    assert np.allclose(premul(left), premul(right), atol=3e-6)


def test_blend_placement_clips():
    rng = np.random.default_rng(5)
    d, s = rgba(rng, 6, 8), rgba(rng, 4, 4, alpha=1.0)
    out = O.blend_pass(d.copy(), s, 6, 4, 3)  # only the top-left 2x2 of `s` lands inside
    assert np.array_equal(out[4:, 6:], s[:2, :2])
    mask = np.ones((6, 8), bool); mask[4:, 6:] = False
    assert np.array_equal(out[mask], d[mask])
    assert np.array_equal(O.blend_pass(d.copy(), s, -10, -10, 3), d)  # completely outside


# ---------------------------------------------------------------- bilinear sampling (AffineSample::BiLinear is rejected, command.rs:1659-1665)
@SET
@given(seed=st.integers(0, 2**31 - 1), w=st.integers(1, 9), h=st.integers(1, 9))
def test_resize_identity_and_constants(seed, w, h):
    rng = np.random.default_rng(seed)
    t = rgba(rng, h, w)
    for sampling in (0, 1):
        assert np.array_equal(O.resize_pass(t, w, h, sampling), t)      # same size: texel centres map onto texel centres
    const = np.empty((h, w, 4), np.float32); const[:] = rng.random(4, dtype=np.float32)
    out = O.resize_pass(const, 2 * w + 1, 3 * h + 2, 1)
    assert np.array_equal(out, np.broadcast_to(const[0, 0], out.shape))  # fma(a, c - c, c) == c: no drift on flat fields


def test_bilinear_is_linear_interpolation_with_edge_clamp():
    ramp = np.zeros((1, 4, 4), np.float32); ramp[0, :, 0] = [0.0, 1.0, 2.0, 3.0]; ramp[..., 3] = 1.0
    out = O.resize_pass(ramp, 8, 1, 1)[0, :, 0]   # centres at (i + 0.5) / 2 - 0.5 = -0.25, 0.25, 0.75, ...
    assert np.allclose(out, [0.0, 0.25, 0.75, 1.25, 1.75, 2.25, 2.75, 3.0])
    near = O.resize_pass(ramp, 8, 1, 0)[0, :, 0]
    assert np.array_equal(near, [0, 0, 1, 1, 2, 2, 3, 3])


@SET
@given(seed=st.integers(0, 2**31 - 1))
def test_affine_identity_and_integer_shift(seed):
    rng = np.random.default_rng(seed)
    src = rgba(rng, 7, 9)
    ident = np.eye(3, dtype=np.float32).reshape(9)
    for sampling in (0, 1):
        dst = np.zeros((7, 9, 4), np.float32)
        O.paint_affine(dst, src, ident, sampling)
        assert np.array_equal(dst, src)
        sx, sy = 2, 3
        inv = np.array([1, 0, -sx, 0, 1, -sy, 0, 0, 1], np.float32)
        bg = rgba(rng, 7, 9)
        dst = bg.copy()
        O.paint_affine(dst, src, inv, sampling)
        assert np.array_equal(dst[sy:, sx:], src[:7 - sy, :9 - sx])   # covered: exact copy
        assert np.array_equal(dst[:sy], bg[:sy]) and np.array_equal(dst[:, :sx], bg[:, :sx])  # uncovered: untouched


# ---------------------------------------------------------------- planar YUV 4:2:0 (program.rs:794-938 lowers Block::Pixel only)
@pytest.mark.parametrize("kr,kb", [(0.2126, 0.0722), (0.299, 0.114), (0.2627, 0.0593)])
@pytest.mark.parametrize("full_range", [False, True])
def test_yuv_grey_axis_and_round_trip(kr, kb, full_range):
    """Neutral chroma decodes to R' = G' = B' = the luma ramp; encode(decode(.)) gives the codes back when every
    2x2 block is flat (4:2:0 carries one chroma sample per block)."""
    h, w = 16, 32
    lo, hi = (0, 255) if full_range else (16, 235)
    y = np.linspace(lo, hi, w * h // 4).round().astype(np.uint8).reshape(h // 2, w // 2).repeat(2, 0).repeat(2, 1)
    u = np.full((h // 2, w // 2), 128, np.uint8); v = u.copy()
    tex = O.decode_yuv420(y, u, v, w, h, kr, kb, full_range, False, 0, O.TR_LINEAR)
    assert np.allclose(tex[..., 0], tex[..., 1], atol=1e-6) and np.allclose(tex[..., 1], tex[..., 2], atol=1e-6)
    assert np.allclose(tex[..., 0], (y.astype(np.float32) - lo) / (hi - lo), atol=1e-6)
    assert np.all(tex[..., 3] == 1.0)
    rng = np.random.default_rng(11)
    c_lo, c_hi = (0, 256) if full_range else (16, 241)
    y2 = rng.integers(lo, hi + 1, (h // 2, w // 2), dtype=np.uint8).repeat(2, 0).repeat(2, 1)
    u2 = rng.integers(c_lo, c_hi, (h // 2, w // 2), dtype=np.uint8); v2 = rng.integers(c_lo, c_hi, (h // 2, w // 2), dtype=np.uint8)
    for tr in (O.TR_LINEAR, O.TR_BT709):
        t2 = O.decode_yuv420(y2, u2, v2, w, h, kr, kb, full_range, False, 0, tr)
        inside = np.all((t2[..., :3] >= 0.0) & (t2[..., :3] <= 1.0), -1)  # out-of-gamut YUV triples do not survive the [0, 1] working range of an EOTF
        ye, ue, ve = O.encode_yuv420(t2, kr, kb, full_range, tr)
        assert np.max(np.abs(ye.astype(int) - y2.astype(int))[inside], initial=0) <= 1
        blk = inside.reshape(h // 2, 2, w // 2, 2).all((1, 3))
        assert np.max(np.abs(ue.astype(int) - u2.astype(int))[blk], initial=0) <= 1
        assert np.max(np.abs(ve.astype(int) - v2.astype(int))[blk], initial=0) <= 1


def test_yuv_nearest_and_bilinear_chroma_agree_on_flat_chroma():
    rng = np.random.default_rng(2)
    h, w = 8, 12
    y = rng.integers(16, 236, (h, w), dtype=np.uint8)
    u = np.full((h // 2, w // 2), 90, np.uint8); v = np.full((h // 2, w // 2), 200, np.uint8)
    a = O.decode_yuv420(y, u, v, w, h, 0.2126, 0.0722, False, False, 0, O.TR_BT709)
    b = O.decode_yuv420(y, u, v, w, h, 0.2126, 0.0722, False, False, 1, O.TR_BT709)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("src_wh,dst_wh", [((64, 48), (32, 24)), ((300, 200), (517, 333)), ((1000, 7), (400, 400)), ((5, 5), (1, 1))])
def test_reference_resize_is_an_8_bit_coordinate_lookup(src_wh, dst_wh):
    """command.rs:1675-1702 + bilinear.frag:14-20 + palette.frag:21-32, restated independently in numpy: the (u, v) ramp
    rests in an f16 texture, is packed into an RGBA8 `Scalars` register by TRUNCATION, unpacked to k / 255 (f16 again), biased
    by half a texel OF THE GRID and used for a nearest fetch -- so at most 256 distinct source columns / rows are ever
    read.  Integer / index work: exact."""
    (sw, sh), (dw, dh) = src_wh, dst_wh
    src = np.random.default_rng(sw * 31 + dw).integers(0, 256, (sh, sw, 4), dtype=np.uint8)
    got = O.resize(O.Image(O.srgb_rgba8(sw, sh), src.reshape(sh, sw * 4)), (dw, dh), "reference").data.reshape(dh, dw, 4)
    f32, f16 = np.float32, lambda a: a.astype(np.float16).astype(np.float32)

    def index(n_dst, n_src):
        u = f16((np.arange(n_dst, dtype=f32) + f32(0.5)) / f32(n_dst))                 # the draw wrote an Rgba16Float attachment
        k = (np.clip(u, f32(0), f32(1)) * f32(255)).astype(np.uint32)                  # mux_uint: truncation
        c = f16(k.astype(f32) / f32(255)) + f32(0.5) / f32(n_dst)                      # demux_uint, then palette.frag's bias
        return np.clip(np.floor(c * f32(n_src)).astype(np.int64), 0, n_src - 1)        # nearest, clamp to edge
    xs, ys = index(dw, sw), index(dh, sh)
    assert np.array_equal(got, src[ys][:, xs])
    assert len(np.unique(xs)) <= 256 and len(np.unique(ys)) <= 256


@SET
@given(seed=st.integers(0, 2**31 - 1), angle=st.floats(-3.1, 3.1), sx=st.floats(0.4, 2.5), sy=st.floats(0.4, 2.5),
       tx=st.floats(-20, 40), ty=st.floats(-20, 40))
def test_affine_nearest_is_the_inverse_map_of_pixel_centres(seed, angle, sx, sy, tx, ty):
    """program.rs:1898-1935 + box.vert:43-59 + copy.frag:8-10, restated in float64: dst pixel (i, j) is covered iff
    p = A^-1 (i + 1/2, j + 1/2) lies inside `above`, and then takes above[floor p_y][floor p_x]; uncovered pixels keep `below`.
    Pixels whose p is within 1e-3 of a texel edge are left out (f32 vs f64 may floor them differently); index work: exact."""
    rng = np.random.default_rng(seed)
    aw, ah, bw, bh = 23, 17, 61, 47
    above, below = rgba(rng, ah, aw), rgba(rng, bh, bw)
    A = O.shift(tx, ty) @ O.rotate(angle) @ O.scale(sx, sy)      # Affine::{scale, rotate, shift} left-multiply
    inv = np.linalg.inv(A.astype(np.float64))
    dst = below.copy()
    O.paint_affine(dst, above, inv.astype(np.float32).reshape(9), 0)
    jj, ii = np.mgrid[0:bh, 0:bw]
    p = np.einsum("rc,chw->rhw", inv, np.stack([ii + 0.5, jj + 0.5, np.ones_like(ii, dtype=np.float64)]))
    px, py = p[0] / p[2], p[1] / p[2]
    near_edge = (np.abs(px - np.round(px)) < 1e-3) | (np.abs(py - np.round(py)) < 1e-3)
    inside = (px >= 0) & (px < aw) & (py >= 0) & (py < ah)
    fx, fy = np.clip(np.floor(px).astype(int), 0, aw - 1), np.clip(np.floor(py).astype(int), 0, ah - 1)
    exp = np.where(inside[..., None], above[fy, fx], below)
    sure = ~near_edge
    assert np.array_equal(dst[sure], exp[sure])


def test_yuv_bt709_colour_bars_known_answers():
    """The 100 % colour bars in 8-bit limited-range BT.709 Y'CbCr (ITU-R BT.709-6 section 3 quantisation; the values every
    test-pattern generator prints): a known-answer test for the planar YUV semantics, which have no reference implementation."""
    bars = {  # R, G, B -> Y, Cb, Cr
        (1, 1, 1): (235, 128, 128), (1, 1, 0): (219, 16, 138), (0, 1, 1): (188, 154, 16), (0, 1, 0): (173, 42, 26),
        (1, 0, 1): (78, 214, 230), (1, 0, 0): (63, 102, 240), (0, 0, 1): (32, 240, 118), (0, 0, 0): (16, 128, 128)}
    for rgb, (Y, Cb, Cr) in bars.items():
        tex = np.zeros((4, 4, 4), np.float32)
        tex[..., :3] = rgb
        tex[..., 3] = 1
        y, u, v = O.encode_yuv420(tex, 0.2126, 0.0722)
        assert (y == Y).all() and (u == Cb).all() and (v == Cr).all(), (rgb, y[0, 0], u[0, 0], v[0, 0])
        back = O.decode_yuv420(np.full((4, 4), Y, np.uint8), np.full((2, 2), Cb, np.uint8), np.full((2, 2), Cr, np.uint8), 4, 4, 0.2126, 0.0722)
        assert np.abs(back[..., :3] - np.array(rgb, np.float32)).max() < 0.012 and (back[..., 3] == 1).all()  # half a code of 219, through the EOTF
