"""command -- the typed op builder, mirroring /root/reference/lib/zosimos/src/command.rs.

`CommandBuffer` has the reference's method names, argument meaning and error behaviour
(`CommandError` with the reference's kinds).  The checks and the parameter preparation run in the
C++ host layer (zosimos_b200/csrc/host.cpp, C view include/zosimos_host.h); this module is the
ctypes veneer the parity tests drive, so they read like lib/zosimos/tests/blend.rs.
"""
from __future__ import annotations

import ctypes as C
import enum
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _ffi
from .buffer import (Block, ByteLayout, Color, ColorChannel, ColorModel, Descriptor, Primaries, SampleBits, SampleParts, Texel,
                     Transfer, Whitepoint, YuvMatrix)


class CommandErrorKind(enum.IntEnum):  # command.rs:3612-3647
    BadDescriptor = 1
    ConflictingTypes = 2
    GenericTypeError = 3
    Other = 4
    Unimplemented = 5
    ConcreteDescriptorRequired = 6


class CommandError(Exception):
    def __init__(self, kind: int, message: str):
        super().__init__("%s: %s" % (CommandErrorKind(kind).name, message))
        self.kind = CommandErrorKind(kind)

    def is_type_err(self) -> bool:  # command.rs:3637-3646
        return self.kind in (CommandErrorKind.GenericTypeError, CommandErrorKind.ConflictingTypes, CommandErrorKind.BadDescriptor)


class ZoshRect(C.Structure):
    _fields_ = [("x", C.c_uint32), ("y", C.c_uint32), ("max_x", C.c_uint32), ("max_y", C.c_uint32)]


_P = C.c_void_p
_F9 = C.c_float * 9
_HOST_SIGNATURES = {
    "zosh_last_error": (C.c_char_p, []),
    "zosh_to_xyz_matrix": (C.c_int32, [C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]),
    "zosh_adaptation_matrix": (C.c_int32, [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]),
    "zosh_whitepoint_xyz": (C.c_int32, [C.c_uint32, C.POINTER(C.c_float)]),
    "zosh_affine_identity": (None, [C.POINTER(C.c_float)]),
    "zosh_affine_scale": (None, [C.POINTER(C.c_float), C.c_float, C.c_float]),
    "zosh_affine_rotate": (None, [C.POINTER(C.c_float), C.c_float]),
    "zosh_affine_shift": (None, [C.POINTER(C.c_float), C.c_float, C.c_float]),
    "zosh_rect_normalize": (ZoshRect, [ZoshRect]),
    "zosh_cb_new": (_P, []),
    "zosh_cb_free": (None, [_P]),
    "zosh_cb_input": (C.c_int32, [_P, C.POINTER(_ffi.ZosDesc), C.POINTER(C.c_int32)]),
    "zosh_cb_output": (C.c_int32, [_P, C.c_int32, C.POINTER(C.c_int32)]),
    "zosh_cb_describe": (C.c_int32, [_P, C.c_int32, C.POINTER(_ffi.ZosDesc)]),
    "zosh_cb_color_convert": (C.c_int32, [_P, C.c_int32, C.POINTER(_ffi.ZosDesc), C.POINTER(C.c_int32)]),
    "zosh_cb_chromatic_adaptation": (C.c_int32, [_P, C.c_int32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32)]),
    "zosh_cb_inscribe": (C.c_int32, [_P, C.c_int32, ZoshRect, C.c_int32, C.POINTER(C.c_int32)]),
    "zosh_cb_crop": (C.c_int32, [_P, C.c_int32, ZoshRect, C.POINTER(C.c_int32)]),
    "zosh_cb_affine": (C.c_int32, [_P, C.c_int32, C.POINTER(C.c_float), C.c_uint32, C.c_int32, C.POINTER(C.c_int32)]),
    "zosh_cb_resize": (C.c_int32, [_P, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32)]),
    "zosh_cb_blend": (C.c_int32, [_P, C.c_int32, ZoshRect, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "zosh_cb_transmute": (C.c_int32, [_P, C.c_int32, C.POINTER(_ffi.ZosDesc), C.POINTER(C.c_int32)]),
    "zosh_cb_bilinear": (C.c_int32, [_P, C.POINTER(_ffi.ZosDesc), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "zosh_cb_solid_rgba": (C.c_int32, [_P, C.POINTER(_ffi.ZosDesc), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "zosh_cb_num_ops": (C.c_uint32, [_P]),
    "zosh_cb_dynamic": (C.c_int32, [_P, C.c_int32, C.c_int32, C.c_char_p, C.POINTER(_ffi.ZosDesc), C.c_void_p, C.c_uint64, C.POINTER(C.c_int32)]),
    "zosh_cb_buffer_init": (C.c_int32, [_P, C.c_void_p, C.c_uint64, C.POINTER(C.c_int32)]),
    "zosh_cb_buffer_zero": (C.c_int32, [_P, C.c_uint64, C.POINTER(C.c_int32)]),
    "zosh_cb_buffer_size": (C.c_int32, [_P, C.c_int32, C.POINTER(C.c_uint64)]),
    "zosh_cb_from_buffer": (C.c_int32, [_P, C.c_int32, C.POINTER(_ffi.ZosDesc), C.POINTER(C.c_int32)]),
    "zosh_cb_with_buffer_bilinear": (C.c_int32, [_P, C.c_int32, C.POINTER(_ffi.ZosDesc), C.POINTER(C.c_int32)]),
    "zosh_cb_distribution_normal2d": (C.c_int32, [_P, C.POINTER(_ffi.ZosDesc), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "zosh_cb_distribution_fractal_noise": (C.c_int32, [_P, C.POINTER(_ffi.ZosDesc), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "zosh_normal2d_with_diagonal": (None, [C.c_float, C.c_float, C.POINTER(C.c_float)]),
    "zosh_normal2d_with_direction": (None, [C.c_float, C.c_float, C.POINTER(C.c_float)]),
    "zosh_fractal_noise_with_octaves": (None, [C.c_uint32, C.POINTER(C.c_float)]),
    "zosh_fractal_noise_set_damping": (None, [C.POINTER(C.c_float), C.c_float]),
    "zosh_cb_derivative": (C.c_int32, [_P, C.c_int32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32)]),
    "zosh_cb_palette": (C.c_int32, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "zosh_cb_extract": (C.c_int32, [_P, C.c_int32, C.c_uint32, C.POINTER(C.c_int32)]),
    "zosh_cb_inject": (C.c_int32, [_P, C.c_int32, C.c_uint32, C.c_int32, C.POINTER(C.c_int32)]),
    "zosh_cb_with_knob": (C.c_int32, [_P]),
    "zosh_cb_generic": (C.c_int32, [_P, C.POINTER(C.c_int32)]),
    "zosh_cb_input_generic": (C.c_int32, [_P, C.c_int32, C.POINTER(C.c_int32)]),
    "zosh_cb_computed_signature": (C.c_int32, [_P, C.POINTER(_P)]),
    "zosh_signature_free": (None, [_P]),
    "zosh_signature_num_generics": (C.c_uint32, [_P]),
    "zosh_signature_num_inputs": (C.c_uint32, [_P]),
    "zosh_signature_num_outputs": (C.c_uint32, [_P]),
    "zosh_cb_function": (C.c_int32, [_P, _P, C.POINTER(C.c_int32)]),
    "zosh_cb_num_functions": (C.c_uint32, [_P]),
    "zosh_cb_invoke": (C.c_int32, [_P, C.c_int32, C.POINTER(_ffi.ZosDesc), C.c_uint32, C.POINTER(C.c_int32), C.c_uint32, C.POINTER(C.c_int32),
                                   C.c_uint32, C.POINTER(C.c_uint32)]),
    "zosh_link": (C.c_int32, [_P, C.POINTER(_ffi.ZosDesc), C.c_uint32, C.POINTER(_P), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                              C.POINTER(_P)]),
    "zosh_program_register": (C.c_int32, [_P, C.c_int32]),
    "zosh_program_knob": (C.c_uint32, [_P, C.c_uint32, C.c_int32]),
    "zosh_compile": (C.c_int32, [_P, C.POINTER(_P)]),
    "zosh_program_free": (None, [_P]),
    "zosh_program_num_ops": (C.c_uint32, [_P]),
    "zosh_program_ops": (C.POINTER(_ffi.ZosOp), [_P]),
    "zosh_program_lower": (C.c_int32, [_P, _P, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
}
_host = None


def host_lib():
    global _host
    if _host is None:
        l = _ffi.lib()
        for name, (res, args) in _HOST_SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _host = l
    return _host


def _check(st: int):
    if st != 0:
        raise CommandError(st, (host_lib().zosh_last_error() or b"").decode())


@dataclass(frozen=True)
class Register:  # command.rs:31-32
    index: int


@dataclass(frozen=True)
class Rectangle:  # command.rs:311-317
    x: int
    y: int
    max_x: int
    max_y: int

    @staticmethod
    def with_width_height(width: int, height: int) -> "Rectangle":
        return Rectangle(0, 0, width, height)

    @staticmethod
    def with_layout(layout: ByteLayout) -> "Rectangle":
        return Rectangle(0, 0, layout.width, layout.height)

    def width(self) -> int:
        return max(self.max_x - self.x, 0)

    def height(self) -> int:
        return max(self.max_y - self.y, 0)

    def contains(self, o: "Rectangle") -> bool:
        return self.x <= o.x and self.y <= o.y and self.width() - (o.x - self.x) >= o.width() and self.height() - (o.y - self.y) >= o.height()

    def normalize(self) -> "Rectangle":
        """With the reference's `max_y = y + width()` (command.rs:3536-3543)."""
        r = host_lib().zosh_rect_normalize(self._ffi())
        return Rectangle(r.x, r.y, r.max_x, r.max_y)

    def meet(self, o: "Rectangle") -> "Rectangle":
        return Rectangle(max(self.x, o.x), max(self.y, o.y), min(self.max_x, o.max_x), min(self.max_y, o.max_y))

    def join(self, o: "Rectangle") -> "Rectangle":
        return Rectangle(min(self.x, o.x), min(self.y, o.y), max(self.max_x, o.max_x), max(self.max_y, o.max_y))

    def _ffi(self) -> ZoshRect:
        return ZoshRect(self.x, self.y, self.max_x, self.max_y)


class AffineSample(enum.IntEnum):  # command.rs:339-352
    Nearest = 0
    BiLinear = 1


class Affine:
    """command.rs:326-337, 3421-3485: row-major homogeneous matrix mapping `above` pixel coordinates to
    `below` pixel coordinates; scale / rotate / shift multiply from the LEFT, in f32."""

    def __init__(self, sampling: AffineSample = AffineSample.Nearest, transformation: Optional[Sequence[float]] = None):
        self.sampling = AffineSample(sampling)
        self._m = _F9()
        if transformation is None:
            host_lib().zosh_affine_identity(self._m)
        else:
            for i, v in enumerate(np.asarray(transformation, dtype=np.float32).reshape(9)):
                self._m[i] = float(v)

    @staticmethod
    def new(sampling: AffineSample) -> "Affine":
        return Affine(sampling)

    @property
    def transformation(self) -> List[float]:
        return list(self._m)

    def _copy(self) -> "Affine":
        return Affine(self.sampling, list(self._m))

    def scale(self, x: float, y: float) -> "Affine":
        a = self._copy(); host_lib().zosh_affine_scale(a._m, x, y); return a

    def rotate(self, rad: float) -> "Affine":
        a = self._copy(); host_lib().zosh_affine_rotate(a._m, rad); return a

    def shift(self, x: float, y: float) -> "Affine":
        a = self._copy(); host_lib().zosh_affine_shift(a._m, x, y); return a


class ChromaticAdaptationMethod(enum.IntEnum):
    BradfordVonKries = 0
    VonKries = 1
    Xyz = 2
    BradfordNonLinear = 3


class Blend(enum.IntEnum):
    """command.rs:319-323 has only `Alpha`; the other Porter-Duff operators are this backend's."""
    Clear = 0
    Src = 1
    Dst = 2
    Alpha = 3  # source-over
    DstOver = 4
    SrcIn = 5
    DstIn = 6
    SrcOut = 7
    DstOut = 8
    SrcAtop = 9
    DstAtop = 10
    Xor = 11


class ResizeMode(enum.IntEnum):
    Reference = 0  # 8-bit coordinate grid + palette, exactly what CommandBuffer::resize lowers to
    Nearest = 1
    Bilinear = 2


class DerivativeMethod(enum.IntEnum):
    Prewitt = 0
    Sobel = 1
    Scharr3 = 2
    Scharr3To4Bit = 3
    Scharr3To8Bit = 4
    Roberts = 5


class Direction(enum.IntEnum):
    Width = 0
    Height = 1


@dataclass(frozen=True)
class Derivative:
    method: DerivativeMethod
    direction: Direction = Direction.Width


@dataclass(frozen=True)
class Bilinear:  # shaders/bilinear.rs:9-32
    u_min: Sequence[float]
    u_max: Sequence[float]
    v_min: Sequence[float]
    v_max: Sequence[float]
    uv_min: Sequence[float] = (0.0, 0.0, 0.0, 0.0)
    uv_max: Sequence[float] = (0.0, 0.0, 0.0, 0.0)

    def flat(self) -> List[float]:
        return [float(x) for v in (self.u_min, self.u_max, self.v_min, self.v_max, self.uv_min, self.uv_max) for x in v]

    def into_std430(self) -> bytes:
        return np.asarray(self.flat(), dtype=np.float32).tobytes()


class DistributionNormal2d:  # shaders/distribution_normal2d.rs:9-100
    def __init__(self, params: Sequence[float]):
        self.params = [float(x) for x in params]  # expectation[2], covariance_inverse[4] row major, pseudo_determinant

    @staticmethod
    def with_diagonal(var0: float, var1: float) -> "DistributionNormal2d":
        out = (C.c_float * 7)()
        host_lib().zosh_normal2d_with_diagonal(var0, var1, out)
        return DistributionNormal2d(list(out))

    @staticmethod
    def with_direction(direction: Sequence[float]) -> "DistributionNormal2d":
        out = (C.c_float * 7)()
        host_lib().zosh_normal2d_with_direction(float(direction[0]), float(direction[1]), out)
        return DistributionNormal2d(list(out))

    def into_std430(self) -> bytes:
        return np.asarray(self.params + [0.0], dtype=np.float32).tobytes()


class FractalNoise:  # shaders/fractal_noise.rs:9-49
    def __init__(self, params: Sequence[float]):
        self.params = [float(x) for x in params]  # initial_scale[2], amplitude, damping, num_octaves

    @staticmethod
    def with_octaves(num_octaves: int) -> "FractalNoise":
        out = (C.c_float * 5)()
        host_lib().zosh_fractal_noise_with_octaves(int(num_octaves), out)
        return FractalNoise(list(out))

    def set_damping(self, damping: float):
        p = (C.c_float * 5)(*self.params)
        host_lib().zosh_fractal_noise_set_damping(p, float(damping))
        self.params = list(p)

    def into_std430(self) -> bytes:
        return np.asarray(self.params[:4], dtype=np.float32).tobytes() + np.asarray([int(self.params[4]), 0], dtype=np.uint32).tobytes()


@dataclass(frozen=True)
class Palette:  # command.rs:566-580
    width: Optional[ColorChannel] = None
    height: Optional[ColorChannel] = None
    width_base: int = 0
    height_base: int = 0


@dataclass(frozen=True)
class RegisterKnob:
    link_idx: int
    register: Register


def descriptor_from_ffi(d: _ffi.ZosDesc) -> Descriptor:
    texel = Texel(Block(d.block), SampleBits(d.bits), SampleParts(d.parts))
    color = Color(ColorModel(d.color), Transfer(d.transfer) if d.transfer < 0x100 else Transfer.Linear, Primaries(d.primaries), Whitepoint(d.whitepoint))
    ts = int(d.texel_stride)
    return Descriptor(ByteLayout(int(d.width), int(d.height), int(d.width) * ts, ts), color, texel, YuvMatrix(d.yuv_matrix),
                      bool(d.yuv_full_range), int(d.chroma_filter))


class CommandBuffer:
    """command.rs:634-705.  Every method pushes one operation and returns its `Register`."""

    def __init__(self):
        self._h = host_lib().zosh_cb_new()
        self._knobs = {}  # register index -> knob id

    def __del__(self):
        try:
            if self._h:
                host_lib().zosh_cb_free(self._h)
                self._h = None
        except Exception:
            pass

    # -- functions and generics (host.cpp: a command buffer that declares a generic records its builder calls, `invoke` replays them)
    _is_template = False

    def generic(self, declaration: "GenericDeclaration" = None) -> "GenericVar":  # command.rs:856-870
        v = C.c_int32()
        _check(host_lib().zosh_cb_generic(self._h, C.byref(v)))
        self._is_template = True
        return GenericVar(int(v.value))

    def input_generic(self, var: "GenericVar") -> Register:  # command.rs:872-884
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_input_generic(self._h, var.index, C.byref(out)), out)

    def computed_signature(self) -> "FunctionSignature":  # command.rs:886-905
        h = _P()
        _check(host_lib().zosh_cb_computed_signature(self._h, C.byref(h)))
        return FunctionSignature(h, self)

    def function(self, signature: "FunctionSignature") -> "FunctionVar":  # command.rs:907-922
        f = C.c_int32()
        _check(host_lib().zosh_cb_function(self._h, signature._h, C.byref(f)))
        return FunctionVar(int(f.value))

    def register_descriptor(self, reg: Register) -> Descriptor:  # the concrete type bound to a generic
        return self.describe_reg(reg)

    def invoke(self, function: "FunctionVar", arguments: "InvocationArguments") -> List[Register]:  # command.rs:2821-2869
        ng, na = len(arguments.generics), len(arguments.arguments)
        def bound(g):  # inside a template a generic of the callee may be bound to one of the template's own generics
            if isinstance(g, GenericVar):
                d = _ffi.ZosDesc()
                d.reserved = 0x80000000 | g.index  # ZOSH_GENERIC_VAR
                return d
            return g.to_ffi()
        gens = (_ffi.ZosDesc * max(ng, 1))(*[bound(g) for g in arguments.generics])
        args = (C.c_int32 * max(na, 1))(*[r.index for r in arguments.arguments])
        results, n = (C.c_int32 * 64)(), C.c_uint32()
        _check(host_lib().zosh_cb_invoke(self._h, function.index, gens, ng, args, na, results, 64, C.byref(n)))
        return [Register(int(results[i])) for i in range(n.value)]

    # -- plumbing
    def _reg(self, st: int, out: C.c_int32) -> Register:
        knob, self._pending_knob = self._pending_knob, 0  # a failed builder must not leave the mark for the next call (the host layer clears its side too)
        _check(st)
        r = Register(int(out.value))
        if knob:
            self._knobs[r.index] = knob
        return r

    _pending_knob = 0

    def describe_reg(self, reg: Register) -> Descriptor:
        d = _ffi.ZosDesc()
        _check(host_lib().zosh_cb_describe(self._h, reg.index, C.byref(d)))
        return descriptor_from_ffi(d)

    def with_knob(self) -> "WithKnob":
        """command.rs:1865-1874: the operation built through the returned wrapper gets a knob -- its parameter block can be
        overridden at run time."""
        return WithKnob(self)

    # -- operations
    def input(self, desc: Descriptor) -> Register:
        out = C.c_int32(); d = desc.to_ffi()
        return self._reg(host_lib().zosh_cb_input(self._h, C.byref(d), C.byref(out)), out)

    def output(self, src: Register) -> Tuple[Register, Descriptor]:
        out = C.c_int32()
        r = self._reg(host_lib().zosh_cb_output(self._h, src.index, C.byref(out)), out)
        return r, (None if self._is_template else self.describe_reg(src))  # a template's types are bound by the caller

    def color_convert(self, src: Register, color: Color, texel: Texel) -> Register:
        out = C.c_int32()
        b = texel.bits.bytes()
        d = Descriptor(ByteLayout(1, 1, b, b), color, texel).to_ffi()
        return self._reg(host_lib().zosh_cb_color_convert(self._h, src.index, C.byref(d), C.byref(out)), out)

    def chromatic_adaptation(self, src: Register, method: ChromaticAdaptationMethod, target: Whitepoint) -> Register:
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_chromatic_adaptation(self._h, src.index, int(method), int(target), C.byref(out)), out)

    def inscribe(self, below: Register, rect: Rectangle, above: Register) -> Register:
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_inscribe(self._h, below.index, rect._ffi(), above.index, C.byref(out)), out)

    def blend(self, below: Register, rect: Rectangle, above: Register, blend: Blend = Blend.Alpha) -> Register:
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_blend(self._h, below.index, rect._ffi(), above.index, int(blend), C.byref(out)), out)

    def crop(self, src: Register, rect: Rectangle) -> Register:
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_crop(self._h, src.index, rect._ffi(), C.byref(out)), out)

    def affine(self, below: Register, affine: Affine, above: Register) -> Register:
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_affine(self._h, below.index, affine._m, int(affine.sampling), above.index, C.byref(out)), out)

    def resize(self, below: Register, upper: Tuple[int, int], mode: ResizeMode = ResizeMode.Reference) -> Register:
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_resize(self._h, below.index, int(upper[0]), int(upper[1]), int(mode), C.byref(out)), out)

    def transmute(self, src: Register, target: Descriptor) -> Register:
        out = C.c_int32(); d = target.to_ffi()
        return self._reg(host_lib().zosh_cb_transmute(self._h, src.index, C.byref(d), C.byref(out)), out)

    def bilinear(self, describe: Descriptor, distribution: Bilinear) -> Register:
        out = C.c_int32(); d = describe.to_ffi()
        p = (C.c_float * 24)(*distribution.flat())
        return self._reg(host_lib().zosh_cb_bilinear(self._h, C.byref(d), p, C.byref(out)), out)

    def distribution_normal2d(self, describe: Descriptor, distribution: "DistributionNormal2d") -> Register:
        out = C.c_int32(); d = describe.to_ffi()
        p = (C.c_float * 7)(*distribution.params)
        return self._reg(host_lib().zosh_cb_distribution_normal2d(self._h, C.byref(d), p, C.byref(out)), out)

    def distribution_fractal_noise(self, describe: Descriptor, distribution: "FractalNoise") -> Register:
        out = C.c_int32(); d = describe.to_ffi()
        p = (C.c_float * 5)(*distribution.params)
        return self._reg(host_lib().zosh_cb_distribution_fractal_noise(self._h, C.byref(d), p, C.byref(out)), out)

    # -- user operators (tests/custom.rs)
    def _dynamic(self, src0: int, src1: int, dynamic: "ShaderCommand") -> Register:
        sd = ShaderData()
        source = dynamic.source()
        desc = dynamic.data(sd)
        out = C.c_int32(); d = desc.to_ffi()
        content = sd.content or b""
        buf = C.create_string_buffer(content, len(content)) if content else None
        st = host_lib().zosh_cb_dynamic(self._h, src0, src1, source.encode(), C.byref(d), C.cast(buf, C.c_void_p) if buf else None,
                                        len(content), C.byref(out))
        return self._reg(st, out)

    def construct_dynamic(self, dynamic: "ShaderCommand") -> Register:  # command.rs:2933-2961
        return self._dynamic(-1, -1, dynamic)

    def unary_dynamic(self, op: Register, dynamic: "ShaderCommand") -> Register:  # command.rs:2963-3002
        return self._dynamic(op.index, -1, dynamic)

    def binary_dynamic(self, lhs: Register, rhs: Register, dynamic: "ShaderCommand") -> Register:  # command.rs:3004-3060
        return self._dynamic(lhs.index, rhs.index, dynamic)

    # -- byte buffers (tests/buffer.rs)
    def buffer_init(self, init: bytes) -> Register:  # command.rs:1777-1790 (with_knob().buffer_init: :1938)
        out = C.c_int32(); data = bytes(init)
        buf = C.create_string_buffer(data, len(data))
        return self._reg(host_lib().zosh_cb_buffer_init(self._h, C.cast(buf, C.c_void_p), len(data), C.byref(out)), out)

    def buffer_zero(self, length: int) -> Register:  # command.rs:1793-1803
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_buffer_zero(self._h, int(length), C.byref(out)), out)

    def buffer_size(self, reg: Register) -> int:
        out = C.c_uint64()
        _check(host_lib().zosh_cb_buffer_size(self._h, reg.index, C.byref(out)))
        return int(out.value)

    def from_buffer(self, buffer_reg: Register, descriptor: Descriptor) -> Register:  # command.rs:937-968
        out = C.c_int32(); d = descriptor.to_ffi()
        return self._reg(host_lib().zosh_cb_from_buffer(self._h, buffer_reg.index, C.byref(d), C.byref(out)), out)

    def with_buffer(self, buffer_reg: Register) -> "WithBuffer":  # command.rs:1876-1890
        self.buffer_size(buffer_reg)  # TYPE_ERR unless it is a buffer register
        return WithBuffer(self, buffer_reg)

    def solid_rgba(self, describe: Descriptor, color: Sequence[float]) -> Register:
        out = C.c_int32(); d = describe.to_ffi()
        c = (C.c_float * 4)(*[float(x) for x in color])
        return self._reg(host_lib().zosh_cb_solid_rgba(self._h, C.byref(d), c, C.byref(out)), out)

    def derivative(self, image: Register, config: Derivative) -> Register:
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_derivative(self._h, image.index, int(config.method), int(config.direction), C.byref(out)), out)

    def palette(self, palette: Register, config: Palette, indices: Register) -> Register:
        pos = {ColorChannel.R: [1, 0, 0, 0], ColorChannel.G: [0, 1, 0, 0], ColorChannel.B: [0, 0, 1, 0]}
        if (config.width is not None and config.width not in pos) or (config.height is not None and config.height not in pos):
            raise CommandError(CommandErrorKind.GenericTypeError, "palette: channel position")  # command.rs:1453-1463
        xc = (C.c_float * 4)(*[float(v) for v in (pos[config.width] if config.width else [0, 0, 0, 0])])
        yc = (C.c_float * 4)(*[float(v) for v in (pos[config.height] if config.height else [0, 0, 0, 0])])
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_palette(self._h, palette.index, indices.index, xc, yc, C.byref(out)), out)

    _CHANNEL = {ColorChannel.R: 0, ColorChannel.G: 1, ColorChannel.B: 2, ColorChannel.Alpha: 3}

    def extract(self, src: Register, channel: ColorChannel) -> Register:
        if channel not in self._CHANNEL:
            raise CommandError(CommandErrorKind.Other, "extract: channel")
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_extract(self._h, src.index, self._CHANNEL[channel], C.byref(out)), out)

    def inject(self, below: Register, channel: ColorChannel, above: Register) -> Register:
        if channel not in self._CHANNEL:
            raise CommandError(CommandErrorKind.Other, "inject: channel")
        out = C.c_int32()
        return self._reg(host_lib().zosh_cb_inject(self._h, below.index, self._CHANNEL[channel], above.index, C.byref(out)), out)


class WithKnob:
    """command.rs:1846-1874 `WithKnob<'lt>`: borrows the command buffer for ONE builder call.  The mark is set when that
    call is made, after its arguments were evaluated, so `cb.with_knob().op(cb.other_op(..))` knobs `op` like in Rust."""

    def __init__(self, cb: "CommandBuffer"):
        self._cb = cb

    def __getattr__(self, name):
        fn = getattr(self._cb, name)
        if name.startswith("_") or not callable(fn):
            raise AttributeError(name)

        def build(*args, **kwargs):
            self._cb._pending_knob = int(host_lib().zosh_cb_with_knob(self._cb._h))
            return fn(*args, **kwargs)
        return build


class ShaderData:  # command/dynamic.rs:40-58
    def __init__(self):
        self.content: Optional[bytes] = None

    def set_data(self, data) -> None:
        self.content = np.asarray(data).tobytes() if not isinstance(data, (bytes, bytearray)) else bytes(data)


class ShaderCommand:  # command/dynamic.rs:7-28; the source is CUDA C instead of SPIR-V (see include/zosimos_cuda.h, zos_dynamic_create)
    def source(self) -> str:
        raise NotImplementedError

    def data(self, data: ShaderData) -> Descriptor:
        raise NotImplementedError


class WithBuffer:  # command.rs:1963-2060: the next operation's parameter block is read from a device buffer
    def __init__(self, cb: "CommandBuffer", buffer_reg: Register):
        self._cb, self._buf = cb, buffer_reg

    def bilinear(self, describe: Descriptor, distribution: Optional[Bilinear] = None) -> Register:
        """The distribution argument only types the call in the reference; the values come from the buffer."""
        out = C.c_int32(); d = describe.to_ffi()
        return self._cb._reg(host_lib().zosh_cb_with_buffer_bilinear(self._cb._h, self._buf.index, C.byref(d), C.byref(out)), out)


# ---- functions and generics (command.rs:856-922, 2083-2185, 2821-2869; tests/generic.rs) -------------------
# A command buffer that declares a generic becomes a TEMPLATE: its operations are recorded, not built, because
# their descriptors depend on the types the caller binds.  `invoke` monomorphises: it replays the callee's
# record into the caller with the generic inputs replaced by the argument registers (the reference does the
# same at link time, command.rs:2083-2185; here the callee travels inside the signature object, and
# `Linker.link` checks that the linked functions are the ones that were invoked).  All of it lives in host.cpp
# (zosh_cb_generic / zosh_cb_invoke / zosh_link); these classes carry the handles.
@dataclass(frozen=True)
class GenericDeclaration:
    bounds: Sequence = ()


@dataclass(frozen=True)
class GenericVar:
    index: int


@dataclass(frozen=True)
class FunctionVar:
    index: int


@dataclass(frozen=True)
class InvocationArguments:
    generics: Sequence  # Descriptor, or (inside a template) a GenericVar of the calling template
    arguments: Sequence[Register]


class FunctionSignature:  # command::CommandSignature
    def __init__(self, handle, template: "CommandBuffer"):
        self._h, self.template = handle, template  # the template stays alive: zosh_link compares its identity
        l = host_lib()
        self.num_generics = int(l.zosh_signature_num_generics(handle))
        self.num_inputs = int(l.zosh_signature_num_inputs(handle))
        self.num_outputs = int(l.zosh_signature_num_outputs(handle))

    def __del__(self):
        try:
            if self._h:
                host_lib().zosh_signature_free(self._h)
                self._h = None
        except Exception:
            pass


class Linker:
    """command.rs:38-41, 2069: the reference's Linker carries the SPIR-V blobs; this one needs nothing
    (the kernels live in libzosimos_cuda.so)."""

    @staticmethod
    def from_included() -> "Linker":
        return Linker()

    def link(self, main: "CommandBuffer", tys: Sequence[Descriptor], functions: Sequence["CommandBuffer"], links: Sequence[Sequence[int]]) -> "Program":
        """command.rs:2083-2185: program 0 is `main`, program k >= 1 is functions[k - 1]; links[p][f] names the
        program that function variable f of program p calls.  Calls were monomorphised by `invoke`, so linking
        verifies the wiring and compiles `main`; `tys` binds the generics of a generic `main` (its registers are
        translated by `Program.register_index`)."""
        if len(links) != 1 + len(functions):
            raise CommandError(4, "link: one link table per program")
        from .program import Program
        flat = [int(t) for table in links for t in table]
        fns = (_P * max(len(functions), 1))(*[f._h for f in functions])
        tab = (C.c_uint32 * max(len(flat), 1))(*flat)
        per = (C.c_uint32 * len(links))(*[len(t) for t in links])
        h = _P()
        bound = (_ffi.ZosDesc * max(len(tys), 1))(*[d.to_ffi() for d in tys])  # types of a generic entry point's generics
        _check(host_lib().zosh_link(main._h, bound, len(tys), fns, len(functions), tab, per, C.byref(h)))
        return Program(h)

    def compile(self, commands: CommandBuffer) -> "Program":
        from .program import Program
        h = _P()
        _check(host_lib().zosh_compile(commands._h, C.byref(h)))
        return Program(h)


def to_xyz_matrix(primaries: Primaries, whitepoint: Whitepoint) -> np.ndarray:
    m = _F9()
    _check(host_lib().zosh_to_xyz_matrix(int(primaries), int(whitepoint), m))
    return np.array(list(m), dtype=np.float32).reshape(3, 3)


def adaptation_matrix(method: ChromaticAdaptationMethod, src: Whitepoint, dst: Whitepoint) -> np.ndarray:
    m = _F9()
    _check(host_lib().zosh_adaptation_matrix(int(method), int(src), int(dst), m))
    return np.array(list(m), dtype=np.float32).reshape(3, 3)
