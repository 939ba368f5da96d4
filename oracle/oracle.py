"""CPU oracle driver -- TEST INFRASTRUCTURE ONLY (see the header of zos_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (zosimos_b200) never does.

This file restates, in numpy + the C passes of zos_oracle.c, how the reference *lowers and runs*
a command buffer: every register lives in memory in its declared texel format, every operand is
decoded into a texture before a draw and every draw result is encoded back
(/root/reference/lib/zosimos/src/program.rs:1475-1478, 1531-1532).  Host-side parameter
preparation follows lib/zosimos/src/command.rs (file:line cited per function).

Third-party arithmetic that is NOT under /root/reference:
  * image-canvas 0.5.1 (Cargo.lock:1739-1740): Primaries::to_xyz_row_matrix(Whitepoint),
    Whitepoint::to_xyz().  Restated from the published construction (chromaticities -> XYZ with
    the white point fixing the channel scales).  PARITY UNPINNED at the last bit; anchored by the
    `adapted`, `oklab`, `srlab2` goldens (BT.709/D65, D50).
  * palette 0.7.6 (Cargo.lock:2311-2312): chromatic_adaptation::TransformMatrix (Bradford,
    VonKries, XyzScaling cone matrices as documented by Lindbloom).  Anchored by `adapted`.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, replace
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libzos_oracle.so")

# ---- numeric codes: stage.frag:107-170 / shaders/stage.rs:74-119 ------------------------------
TR_BT709, TR_BT470M, TR_BT601, TR_SMPTE240, TR_LINEAR, TR_SRGB = 0, 1, 2, 3, 4, 5
TR_BT2020_10, TR_BT2020_12, TR_SMPTE2084, TR_BT2100PQ, TR_BT2100HLG, TR_LINEAR_SCENE = 6, 7, 8, 9, 10, 11
TR_LABLCH = 0x100
P_A, P_R, P_G, P_B, P_LUMA, P_LUMAA, P_RGB, P_BGR, P_RGBA = 0, 1, 2, 3, 4, 5, 6, 7, 8
P_BGRA, P_ARGB, P_ABGR, P_YUV, P_LAB, P_LABA, P_LCH, P_LCHA = 10, 12, 14, 16, 17, 18, 19, 20
B_UINT8, B_UINT332, B_UINT233, B_UINT16, B_UINT4X4, B_UINT565, B_UINT8X2, B_UINT8X3 = 0, 1, 2, 3, 4, 7, 8, 9
B_UINT8X4, B_UINT16X2, B_UINT16X3, B_UINT16X4 = 10, 11, 12, 13
B_UINT2101010, B_UINT1010102 = 14, 15  # image-canvas names (stage.rs:112-113); 15 = standard RGB10A2
B_FLOAT16X4, B_FLOAT32X4 = 18, 19
ST_STAGED, ST_SRGB8, ST_UNORM8, ST_FLOAT = 0, 1, 2, 3


def build(force: bool = False) -> str:
    """Compile the C passes (gcc, a second or two)."""
    src = os.path.join(_HERE, "zos_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Fmt(C.Structure):
    _fields_ = [("transfer", C.c_uint32), ("parts", C.c_uint32), ("bits", C.c_uint32), ("storage", C.c_uint32)]


class Yuv(C.Structure):
    _fields_ = [("kr", C.c_float), ("kb", C.c_float), ("full_range", C.c_uint32), ("nv12", C.c_uint32),
                ("chroma_filter", C.c_uint32), ("transfer", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _bp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# ---- descriptors (image-canvas vocabulary, buffer.rs:14-30) -----------------------------------
@dataclass(frozen=True)
class Texel:
    bits: int
    parts: int

    @property
    def bytes(self) -> int:
        return lib().zo_bits_bytes(self.bits)


PRIMARIES = {  # CIE xy of R, G, B (ITU-R BT.601/709/2020, SMPTE 240M)
    "bt709": ((0.64, 0.33), (0.30, 0.60), (0.15, 0.06)),
    "bt601_525": ((0.630, 0.340), (0.310, 0.595), (0.155, 0.070)),
    "bt601_625": ((0.64, 0.33), (0.29, 0.60), (0.15, 0.06)),
    "smpte240": ((0.630, 0.340), (0.310, 0.595), (0.155, 0.070)),
    "bt2020": ((0.708, 0.292), (0.170, 0.797), (0.131, 0.046)),
    "bt2100": ((0.708, 0.292), (0.170, 0.797), (0.131, 0.046)),
}
WHITEPOINTS = {  # XYZ, Y = 1 (ASTM E308, 2 degree observer)
    "A": (1.09850, 1.0, 0.35585), "B": (0.99072, 1.0, 0.85223), "C": (0.98074, 1.0, 1.18232),
    "D50": (0.96422, 1.0, 0.82521), "D55": (0.95682, 1.0, 0.92149), "D65": (0.95047, 1.0, 1.08883),
    "D75": (0.94972, 1.0, 1.22638), "E": (1.0, 1.0, 1.0), "F2": (0.99186, 1.0, 0.67393),
    "F7": (0.95041, 1.0, 1.08747), "F11": (1.00962, 1.0, 0.64350),
}


@dataclass(frozen=True)
class Color:
    model: str  # "rgb" | "scalars" | "oklab" | "srlab2"
    transfer: int = TR_LINEAR
    primaries: str = "bt709"
    whitepoint: str = "D65"


SRGB = Color("rgb", TR_SRGB, "bt709", "D65")
BT709_RGB = Color("rgb", TR_BT709, "bt709", "D65")
SCALARS_LINEAR = Color("scalars", TR_LINEAR)
OKLAB = Color("oklab")


@dataclass(frozen=True)
class Desc:
    width: int
    height: int
    texel: Texel
    color: Color

    @property
    def row_bytes(self) -> int:
        return self.width * self.texel.bytes


RGBA8 = Texel(B_UINT8X4, P_RGBA)


def srgb_rgba8(w: int, h: int) -> Desc:
    return Desc(w, h, RGBA8, SRGB)


def storage_fmt(d: Desc) -> Fmt:
    """ImageDescriptor::new, program.rs:781-946: native texture format or staging parameters."""
    t, c = d.texel, d.color
    if t.bits == B_UINT8X4 and t.parts in (P_RGBA, P_BGRA) and c.model == "rgb" and c.transfer in (TR_SRGB, TR_LINEAR):
        return Fmt(c.transfer, t.parts, t.bits, ST_SRGB8 if c.transfer == TR_SRGB else ST_UNORM8)
    if c.model in ("rgb", "scalars"):
        if t.bits in (B_FLOAT16X4, B_FLOAT32X4):
            return Fmt(c.transfer, t.parts, t.bits, ST_FLOAT)  # OURS (reference: todo!())
        if t.bytes not in (1, 2, 4) and t.bits != B_UINT16X4:  # UInt16x4: OURS (declared, never defined in stage.frag)
            raise ValueError("unsupported staged texel (stage.rs:63-72)")
        return Fmt(c.transfer, t.parts, t.bits, ST_STAGED)
    if c.model in ("oklab", "srlab2") and t.parts in (P_LCHA, P_LABA):
        tr = TR_LABLCH if t.parts == P_LCHA else TR_LINEAR
        if t.bits in (B_FLOAT16X4, B_FLOAT32X4):
            return Fmt(tr, P_LCHA, t.bits, ST_FLOAT)
        return Fmt(tr, P_LCHA, t.bits, ST_STAGED)  # parts forced to LchA (program.rs:889, 923)
    raise ValueError("no texture representation for %r" % (d,))


@dataclass
class Image:
    """A register: bytes in the declared texel format, tight rows (h, w*bytes)."""
    desc: Desc
    data: np.ndarray

    def __post_init__(self):
        self.data = np.ascontiguousarray(self.data, dtype=np.uint8).reshape(self.desc.height, self.desc.row_bytes)


# ---- passes -----------------------------------------------------------------------------------
def decode(img: Image) -> np.ndarray:
    d = img.desc
    tex = np.empty((d.height, d.width, 4), np.float32)
    f = storage_fmt(d)
    lib().zo_decode(C.byref(f), _bp(img.data), C.c_size_t(d.row_bytes), d.width, d.height, _fp(tex))
    return tex


def encode(desc: Desc, tex: np.ndarray) -> Image:
    tex = _f32(tex)
    assert tex.shape == (desc.height, desc.width, 4), (tex.shape, desc)
    out = np.zeros((desc.height, desc.row_bytes), np.uint8)
    f = storage_fmt(desc)
    lib().zo_encode(C.byref(f), _fp(tex), desc.width, desc.height, _bp(out), C.c_size_t(desc.row_bytes))
    return Image(desc, out)


def _mat(m) -> np.ndarray:
    return _f32(np.asarray(m, dtype=np.float64).reshape(9))


def linear(tex, M):
    tex = _f32(tex); out = np.empty_like(tex); m = _mat(M)
    lib().zo_linear(_fp(m), _fp(tex), _fp(out), C.c_size_t(tex.size // 4))
    return out


def oklab_encode(tex, T):
    tex = _f32(tex); out = np.empty_like(tex); m = _mat(T)
    lib().zo_oklab_encode(_fp(m), _fp(tex), _fp(out), C.c_size_t(tex.size // 4))
    return out


def oklab_decode(tex, T):
    tex = _f32(tex); out = np.empty_like(tex); m = _mat(T)
    lib().zo_oklab_decode(_fp(m), _fp(tex), _fp(out), C.c_size_t(tex.size // 4))
    return out


def srlab2_encode(tex, T):
    tex = _f32(tex); out = np.empty_like(tex); m = _mat(T)
    lib().zo_srlab2_encode(_fp(m), _fp(tex), _fp(out), C.c_size_t(tex.size // 4))
    return out


def srlab2_decode(tex, T, wp_xyz):
    tex = _f32(tex); out = np.empty_like(tex); m = _mat(T); wp = _f32(wp_xyz)
    lib().zo_srlab2_decode(_fp(m), _fp(wp), _fp(tex), _fp(out), C.c_size_t(tex.size // 4))
    return out


def paint_rect(dst_tex, src_tex, sel, tgt):
    """dst_tex is modified in place (Target::Load)."""
    src_tex = _f32(src_tex)
    s = (C.c_int * 4)(*sel); t = (C.c_int * 4)(*tgt)
    lib().zo_paint_rect(_fp(src_tex), src_tex.shape[1], src_tex.shape[0], s, _fp(dst_tex), dst_tex.shape[1],
                        dst_tex.shape[0], t)
    return dst_tex


def paint_affine(dst_tex, src_tex, inv, sampling=0):
    src_tex = _f32(src_tex); m = _f32(inv)
    lib().zo_paint_affine(_fp(src_tex), src_tex.shape[1], src_tex.shape[0], _fp(m), int(sampling), _fp(dst_tex),
                          dst_tex.shape[1], dst_tex.shape[0])
    return dst_tex


def paint_affine_window(dst_rows, dst_y0, src_rows, src_y0, src_full_h, inv, sampling=0):
    """Row-band form: dst_rows = rows [dst_y0, ...) of the destination, src_rows = rows [src_y0, ...) of the
    full source (height src_full_h); coordinates are those of the full images."""
    src_rows = _f32(src_rows); m = _f32(inv)
    lib().zo_paint_affine_window(_fp(src_rows), src_rows.shape[1], int(src_full_h), int(src_y0), _fp(m), int(sampling), _fp(dst_rows),
                                 dst_rows.shape[1], dst_rows.shape[0], int(dst_y0))
    return dst_rows


def gen_bilinear(params: Sequence[Sequence[float]], w: int, h: int):
    p = _f32(np.asarray(params, dtype=np.float32).reshape(24))
    out = np.empty((h, w, 4), np.float32)
    lib().zo_gen_bilinear(_fp(p), _fp(out), w, h)
    return out


def gen_normal2d(params: Sequence[float], w: int, h: int):
    p = _f32(np.asarray(params, dtype=np.float32).reshape(7))
    out = np.empty((h, w, 4), np.float32)
    lib().zo_gen_normal2d(_fp(p), _fp(out), w, h)
    return out


def gen_fractal_noise(params: Sequence[float], w: int, h: int):
    p = _f32(np.asarray(params, dtype=np.float32).reshape(5))
    out = np.empty((h, w, 4), np.float32)
    lib().zo_gen_fractal_noise(_fp(p), _fp(out), w, h)
    return out


def palette_pass(lhs_tex, rhs_tex, x_coord, y_coord):
    lhs_tex = _f32(lhs_tex); rhs_tex = _f32(rhs_tex)
    h, w = rhs_tex.shape[:2]
    out = np.empty((h, w, 4), np.float32)
    xc = _f32(x_coord); yc = _f32(y_coord)
    lib().zo_palette(_fp(lhs_tex), lhs_tex.shape[1], lhs_tex.shape[0], _fp(rhs_tex), w, h, _fp(xc), _fp(yc), _fp(out))
    return out


def resize_pass(src_tex, w, h, sampling):
    src_tex = _f32(src_tex)
    out = np.empty((h, w, 4), np.float32)
    lib().zo_resize(_fp(src_tex), src_tex.shape[1], src_tex.shape[0], _fp(out), w, h, int(sampling))
    return out


def inject_pass(bg, fg, mixv, color):
    bg = _f32(bg); fg = _f32(fg); out = np.empty_like(bg)
    m = _f32(mixv); c = _f32(color)
    lib().zo_inject(_fp(bg), _fp(fg), _fp(m), _fp(c), _fp(out), C.c_size_t(bg.size // 4))
    return out


def box3_pass(tex, M):
    tex = _f32(tex); out = np.empty_like(tex); m = _mat(M)
    lib().zo_box3(_fp(m), _fp(tex), _fp(out), tex.shape[1], tex.shape[0])
    return out


def blend_pass(dst_tex, src_tex, tx, ty, mode):
    src_tex = _f32(src_tex)
    lib().zo_blend(_fp(src_tex), src_tex.shape[1], src_tex.shape[0], _fp(dst_tex), dst_tex.shape[1], dst_tex.shape[0],
                   int(tx), int(ty), int(mode))
    return dst_tex


# ---- host-side parameter preparation ----------------------------------------------------------
def inv3(m):
    """Adjugate inverse in float64, fixed evaluation order (mirrored by the product's host code)."""
    a, b, c, d, e, f, g, h, i = [float(x) for x in np.asarray(m, dtype=np.float64).reshape(9)]
    A = e * i - f * h; B = -(d * i - f * g); Cc = d * h - e * g
    det = a * A + b * B + c * Cc
    return np.array([[A / det, -(b * i - c * h) / det, (b * f - c * e) / det],
                     [B / det, (a * i - c * g) / det, -(a * f - c * d) / det],
                     [Cc / det, -(a * h - b * g) / det, (a * e - b * d) / det]], dtype=np.float64)


def mul3(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(3, 3); b = np.asarray(b, dtype=np.float64).reshape(3, 3)
    o = np.zeros((3, 3))
    for r in range(3):
        for c in range(3):
            o[r, c] = a[r, 0] * b[0, c] + a[r, 1] * b[1, c] + a[r, 2] * b[2, c]
    return o


def to_xyz(primaries: str, whitepoint: str) -> np.ndarray:
    """image-canvas Primaries::to_xyz_row_matrix (call sites command.rs:1023-1073): columns are the
    XYZ of the primaries scaled so that RGB=(1,1,1) maps to the white point."""
    (xr, yr), (xg, yg), (xb, yb) = PRIMARIES[primaries]
    P = np.array([[xr / yr, xg / yg, xb / yb], [1.0, 1.0, 1.0],
                  [(1 - xr - yr) / yr, (1 - xg - yg) / yg, (1 - xb - yb) / yb]], dtype=np.float64)
    w = np.array(WHITEPOINTS[whitepoint], dtype=np.float64)
    Pi = inv3(P)
    S = [Pi[r, 0] * w[0] + Pi[r, 1] * w[1] + Pi[r, 2] * w[2] for r in range(3)]
    return np.array([[P[r, c] * S[c] for c in range(3)] for r in range(3)], dtype=np.float64)


CONE = {
    "bradford": [[0.8951, 0.2664, -0.1614], [-0.7502, 1.7135, 0.0367], [0.0389, -0.0685, 1.0296]],
    "vonkries": [[0.40024, 0.7076, -0.08081], [-0.2263, 1.16532, 0.0457], [0.0, 0.0, 0.91822]],
    "xyz": [[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]],
}


def adaptation_matrix(method: str, src_wp: str, dst_wp: str) -> np.ndarray:
    """palette TransformMatrix::generate_transform_matrix (command.rs:3281-3338)."""
    Mc = np.array(CONE[method], dtype=np.float64)
    s = np.array(WHITEPOINTS[src_wp]); d = np.array(WHITEPOINTS[dst_wp])
    cs = [Mc[r, 0] * s[0] + Mc[r, 1] * s[1] + Mc[r, 2] * s[2] for r in range(3)]
    cd = [Mc[r, 0] * d[0] + Mc[r, 1] * d[1] + Mc[r, 2] * d[2] for r in range(3)]
    D = np.diag([cd[k] / cs[k] for k in range(3)])
    return mul3(inv3(Mc), mul3(D, Mc))


def rotate(r: float) -> np.ndarray:
    """Affine::rotate, command.rs:3451-3466 (f32 cos/sin like the reference)."""
    c, s = np.float32(math.cos(np.float32(r))), np.float32(math.sin(np.float32(r)))
    return np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]], dtype=np.float64)


def shift(x: float, y: float) -> np.ndarray:
    return np.array([[1, 0, x], [0, 1, y], [0, 0, 1]], dtype=np.float64)


def scale(x: float, y: float) -> np.ndarray:
    return np.diag([x, y, 1.0])


# ---- operations, as the reference lowers them (command.rs:2345-2880) --------------------------
def color_convert(src: Image, color: Color, texel: Texel, exact_quirks: bool = True) -> Image:
    """CommandBuffer::color_convert (command.rs:986-1107) + ColorConversion::to_shader (3221-3276)."""
    s = src.desc.color
    dst = Desc(src.desc.width, src.desc.height, texel, color)
    tex = decode(src)
    if s.model == "rgb" and color.model == "rgb":
        if s.whitepoint != color.whitepoint:
            raise ValueError("No conversion")
        if exact_quirks:  # command.rs:1022-1025 + 3227-3228: to_xyz(dst) * inv(to_xyz(src))
            M = mul3(to_xyz(color.primaries, color.whitepoint), inv3(to_xyz(s.primaries, s.whitepoint)))
        else:
            M = mul3(inv3(to_xyz(color.primaries, color.whitepoint)), to_xyz(s.primaries, s.whitepoint))
        out = linear(tex, M)
    elif s.model == "rgb" and color.model == "oklab":
        if s.whitepoint != "D65":
            raise ValueError("No conversion")
        out = oklab_encode(tex, to_xyz(s.primaries, "D65"))
    elif s.model == "oklab" and color.model == "rgb":
        if color.whitepoint != "D65":
            raise ValueError("No conversion")
        out = oklab_decode(tex, inv3(to_xyz(color.primaries, "D65")))
    elif s.model == "rgb" and color.model == "srlab2":
        out = srlab2_encode(tex, to_xyz(s.primaries, s.whitepoint))
    elif s.model == "srlab2" and color.model == "rgb":
        out = srlab2_decode(tex, inv3(to_xyz(color.primaries, color.whitepoint)), WHITEPOINTS[s.whitepoint])
    else:
        raise ValueError("No conversion")
    return encode(dst, out)


def chromatic_adaptation(src: Image, method: str, target: str) -> Image:
    """command.rs:1112-1174, 2527-2548."""
    c = src.desc.color
    if c.model != "rgb":
        raise ValueError("non-rgb chromatic adaptation")
    M = mul3(inv3(to_xyz(c.primaries, target)), mul3(adaptation_matrix(method, c.whitepoint, target),
                                                      to_xyz(c.primaries, c.whitepoint)))
    dst = replace(src.desc, color=replace(c, whitepoint=target))
    return encode(dst, linear(decode(src), M))


def inscribe(below: Image, rect: Tuple[int, int, int, int], above: Image, exact_quirks: bool = True) -> Image:
    """command.rs:1177-1218, 2706-2740.  rect = (x, y, max_x, max_y).  With exact_quirks the placement is
    Rectangle::normalize()'d, whose max_y = y + width() (command.rs:3536-3543)."""
    x, y, mx, my = rect
    aw, ah = above.desc.width, above.desc.height
    if (x, y, mx, my) != (0, 0, aw, ah):
        raise ValueError("inscribe: rect must equal the layout of `above` (command.rs:1196-1198)")
    if aw > below.desc.width or ah > below.desc.height:
        raise ValueError("inscribe: not contained (command.rs:1202-1206)")
    tw = mx - x
    th = (mx - x) if exact_quirks else (my - y)
    tex = decode(below).copy()
    paint_rect(tex, decode(above), (0, 0, aw, ah), (x, y, tw, th))
    return encode(below.desc, tex)


def crop(src: Image, rect: Tuple[int, int, int, int]) -> Image:
    """command.rs:971-978, 2507-2526: the selection is stretched over a target normalised by itself,
    i.e. over the whole (source-sized) output."""
    x, y, mx, my = rect
    tex = np.zeros((src.desc.height, src.desc.width, 4), np.float32)
    tex[..., 2] = 1.0; tex[..., 3] = 1.0  # Target::Discard clears to (0,0,1,1) (program.rs:1494-1506)
    paint_rect(tex, decode(src), (x, y, mx - x, my - y), (0, 0, src.desc.width, src.desc.height))
    return encode(src.desc, tex)


def affine(below: Image, matrix, above: Image, sampling: int = 0) -> Image:
    """command.rs:1636-1673, 2642-2678.  `matrix` maps above's pixel coordinates to below's."""
    m = np.asarray(matrix, dtype=np.float32).astype(np.float64).reshape(3, 3)
    inv = inv3(m)
    tex = decode(below).copy()
    paint_affine(tex, decode(above), inv.astype(np.float32).reshape(9), sampling)
    return encode(below.desc, tex)


def bilinear(desc: Desc, params) -> Image:
    """command.rs:1615-1633; params = (u_min, u_max, v_min, v_max, uv_min, uv_max)."""
    return encode(desc, gen_bilinear(params, desc.width, desc.height))


def normal2d_with_diagonal(var0: float, var1: float):
    """shaders/distribution_normal2d.rs:25-39 -> (expectation[2], covariance_inverse row major[4], pseudo_determinant)."""
    f = np.float32
    var0, var1 = f(var0), f(var1)
    d0 = f(0) if var0 == 0 else f(1) / var0
    d1 = f(0) if var1 == 0 else f(1) / var1
    pi = f(np.pi)
    f0 = f(1) if var0 == 0 else f(2) * pi * var0
    f1 = f(1) if var1 == 0 else f(2) * pi * var1
    return [0.0, 0.0, float(d0), 0.0, 0.0, float(d1), float(f0 * f1)]


def normal2d_with_direction(x: float, y: float):
    """shaders/distribution_normal2d.rs:50-100 (the 'Herbie' forms, in f32)."""
    f = np.float32
    x, y = f(x), f(y)
    length_sq = f(np.float64(x) * np.float64(x) + np.float64(y) * np.float64(y))

    def hyp(a, b):
        return f(np.hypot(np.float64(a), np.float64(b)))

    def sym(a, b):
        h = hyp(a, b)
        return ((f(1) / h) * (a / h)) / (a + b * (b / a))

    def asym(a, b):
        a, b = min(a, b), max(a, b)
        h = hyp(a, b)
        inner = f(np.float64(a) * np.float64(a / b) + np.float64(b))  # mul_add
        return ((a / h) / inner) / h

    return [0.0, 0.0, float(sym(x, x)), float(asym(x, y)), float(asym(y, x)), float(sym(y, y)), float(f(f(f(2.0) * f(np.pi)) * length_sq))]  # 2.0 * PIf32 * length_sq (:96)


def fractal_noise_with_octaves(n: int, damping: float = None):
    """shaders/fractal_noise.rs:21-49 -> (scale.x, scale.y, amplitude, damping, octaves)."""
    f = np.float32
    amp, damp = f(1.0 / float(n)), f(1.0)
    if damping is not None:
        damp = f(damping)
        total = f(1) - f(np.power(damp, f(n)))
        amp = f(1) if abs(total) < 1e-7 else (f(1) - damp) / total
    return [100.0, 100.0, float(amp), float(damp), float(n)]


def distribution_normal2d(desc: Desc, params) -> Image:
    """command.rs distribution_normal2d: a generator like `bilinear`, then the target texel's encode."""
    return encode(desc, gen_normal2d(params, desc.width, desc.height))


def distribution_fractal_noise(desc: Desc, params) -> Image:
    return encode(desc, gen_fractal_noise(params, desc.width, desc.height))


def palette(pal: Image, indices: Image, x_coord, y_coord) -> Image:
    """command.rs:1442-1485, 2741-2759."""
    dst = Desc(indices.desc.width, indices.desc.height, pal.desc.texel, pal.desc.color)
    return encode(dst, palette_pass(decode(pal), decode(indices), x_coord, y_coord))


def resize(below: Image, size: Tuple[int, int], mode: str = "reference") -> Image:
    """command.rs:1675-1702.  mode "reference": RGBA8 linear-scalars coordinate grid + palette lookup
    (coordinates truncated to 8 bits!).  "nearest"/"bilinear": OURS, exact resampling."""
    w, h = size
    if mode == "reference":
        grid = bilinear(Desc(w, h, RGBA8, SCALARS_LINEAR),
                        ([0, 0, 0, 1], [1, 0, 0, 1], [0, 0, 0, 1], [0, 1, 0, 1], [0, 0, 0, 1], [0, 0, 0, 1]))
        return palette(below, grid, [1, 0, 0, 0], [0, 1, 0, 0])
    dst = replace(below.desc, width=w, height=h)
    return encode(dst, resize_pass(decode(below), w, h, 0 if mode == "nearest" else 1))


def transmute(src: Image, desc: Desc) -> Image:
    """command.rs:1276-1350 -> High::Copy: the bytes are reinterpreted."""
    if src.desc.texel.bytes != desc.texel.bytes or (src.desc.width, src.desc.height) != (desc.width, desc.height):
        raise ValueError("invalid transmute")
    return Image(desc, src.data.copy())


def solid(desc: Desc, color) -> Image:
    tex = np.empty((desc.height, desc.width, 4), np.float32)
    tex[:] = np.asarray(color, dtype=np.float32)
    return encode(desc, tex)


def derivative(src: Image, smooth: Sequence[float], direction: str = "width") -> Image:
    """command.rs:3343-3418: weight(dx,dy) = smooth[dy+1] * (+1/2, 0, -1/2)[dx+1] for Direction::Width."""
    M = np.outer(np.asarray(smooth, dtype=np.float32), np.asarray([0.5, 0.0, -0.5], dtype=np.float32))
    if direction != "width":
        M = M.T
    return encode(src.desc, box3_pass(decode(src), M))


def channel_texel(texel: Texel, channel: str) -> Texel:
    return Texel({B_UINT8X4: B_UINT8, B_UINT16X4: B_UINT16}.get(texel.bits, texel.bits), {"R": P_R, "G": P_G, "B": P_B, "A": P_A}[channel])


def extract(src: Image, channel: str) -> Image:
    """command.rs:1221-1266, 2578-2597: full copy; the channel is picked by the destination's parts."""
    dst = Desc(src.desc.width, src.desc.height, channel_texel(src.desc.texel, channel), src.desc.color)
    return encode(dst, decode(src))


def inject(below: Image, channel: str, above: Image) -> Image:
    """command.rs:1360-1440, 2679-2705."""
    mixv = {"R": [1, 0, 0, 0], "G": [0, 1, 0, 0], "B": [0, 0, 1, 0]}[channel]
    color = {P_R: [1, 0, 0, 0], P_G: [0, 1, 0, 0], P_B: [0, 0, 1, 0], P_A: [0, 0, 0, 1], P_LUMA: [1, 0, 0, 0]}[above.desc.texel.parts]
    return encode(below.desc, inject_pass(decode(below), decode(above), mixv, color))


def blend(below: Image, rect: Tuple[int, int, int, int], above: Image, mode: int = 3) -> Image:
    """OURS (command.rs:1510-1519 is UNIMPLEMENTED in the reference): Porter-Duff in linear light,
    straight alpha, `above` placed at rect without scaling."""
    x, y, mx, my = rect
    if (mx - x, my - y) != (above.desc.width, above.desc.height):
        raise ValueError("blend: rect must have the size of `above`")
    tex = decode(below).copy()
    blend_pass(tex, decode(above), x, y, mode)
    return encode(below.desc, tex)


def decode_yuv420(y, u, v, w, h, kr, kb, full_range=False, nv12=False, chroma_filter=0, transfer=TR_BT709):
    y = np.ascontiguousarray(y, np.uint8); u = np.ascontiguousarray(u, np.uint8); v = np.ascontiguousarray(v, np.uint8)
    p = Yuv(kr, kb, int(full_range), int(nv12), int(chroma_filter), int(transfer))
    tex = np.empty((h, w, 4), np.float32)
    lib().zo_decode_yuv420(C.byref(p), _bp(y), C.c_size_t(y.strides[0]), _bp(u), _bp(v), C.c_size_t(u.strides[0]), w, h, _fp(tex))
    return tex


def encode_yuv420(tex, kr, kb, full_range=False, transfer=TR_BT709):
    tex = _f32(tex); h, w = tex.shape[:2]
    cw, ch = (w + 1) // 2, (h + 1) // 2
    y = np.zeros((h, w), np.uint8); u = np.zeros((ch, cw), np.uint8); v = np.zeros((ch, cw), np.uint8)
    p = Yuv(kr, kb, int(full_range), 0, 0, int(transfer))
    lib().zo_encode_yuv420(C.byref(p), _fp(tex), w, h, _bp(y), C.c_size_t(w), _bp(u), _bp(v), C.c_size_t(cw))
    return y, u, v


# ---- blockhash256 (the `blockhash` crate as used by tests/util.rs:21-60) ----------------------
def blockhash256(rgba: np.ndarray) -> str:
    """rgba: (h, w, 4) uint8 with w, h multiples of 16.  Per-block sum of r+g+b (alpha 0 counts as
    white), 4 horizontal bands, bit = value > band median (ties to 1 when the median is bright)."""
    a = np.asarray(rgba).astype(np.int64)
    h, w, _ = a.shape
    bits = 16
    if w < bits and bits % w == 0 and h < bits and bits % h == 0:
        # fewer pixels than blocks: a block is a fraction of ONE pixel (blockhash's weighted method), i.e. pixel replication
        a = np.repeat(np.repeat(a, bits // h, axis=0), bits // w, axis=1)
        h, w, _ = a.shape
    assert w % bits == 0 and h % bits == 0
    v = np.where(a[..., 3] == 0, 765, a[..., 0] + a[..., 1] + a[..., 2])
    bw, bh = w // bits, h // bits
    blocks = v.reshape(bits, bh, bits, bw).sum(axis=(1, 3)).reshape(-1).tolist()
    band = len(blocks) // 4
    maxv = bw * bh * 765
    out = []
    for i in range(4):
        seg = blocks[i * band:(i + 1) * band]
        s = sorted(seg); n = len(s)
        m = s[n // 2] if n % 2 else (s[n // 2 - 1] + s[n // 2]) / 2.0
        for x in seg:
            out.append(1 if (x > m or (abs(x - m) < 1 and m > maxv / 2)) else 0)
    return "%064x" % int("".join(map(str, out)), 2)
