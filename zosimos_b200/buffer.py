"""Descriptors: the image-canvas vocabulary re-exported by the reference's buffer module
(/root/reference/lib/zosimos/src/buffer.rs:2-30): Descriptor {layout: ByteLayout, color: Color,
texel: Texel}, SampleBits, SampleParts, Transfer, Whitepoint, Block.  Numeric values are the codes
the reference feeds its shaders (lib/zosimos/src/shaders/stage.rs:74-119), which are also the
codes of include/zosimos_cuda.h.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, replace
from typing import Optional

from . import _ffi


class Transfer(enum.IntEnum):
    Bt709 = 0
    Bt470M = 1
    Bt601 = 2
    Smpte240 = 3
    Linear = 4
    Srgb = 5
    Bt2020_10bit = 6
    Bt2020_12bit = 7
    Smpte2084 = 8
    Bt2100Pq = 9
    Bt2100Hlg = 10
    LinearScene = 11


class SampleParts(enum.IntEnum):
    A = 0
    R = 1
    G = 2
    B = 3
    Luma = 4
    LumaA = 5
    Rgb = 6
    Bgr = 7
    RgbA = 8
    BgrA = 10
    ARgb = 12
    ABgr = 14
    Yuv = 16
    Lab = 17
    LabA = 18
    Lch = 19
    LchA = 20

    def num_components(self) -> int:
        return {0: 1, 1: 1, 2: 1, 3: 1, 4: 1, 5: 2, 6: 3, 7: 3, 8: 4, 10: 4, 12: 4, 14: 4, 16: 3, 17: 3, 18: 4, 19: 3, 20: 4}[int(self)]


class SampleBits(enum.IntEnum):
    UInt8 = 0
    UInt332 = 1
    UInt233 = 2
    UInt16 = 3
    UInt4x4 = 4
    UInt565 = 7
    UInt8x2 = 8
    UInt8x3 = 9
    UInt8x4 = 10
    UInt16x2 = 11
    UInt16x3 = 12
    UInt16x4 = 13
    UInt2101010 = 14
    UInt1010102 = 15
    Float16x4 = 18
    Float32x4 = 19

    def bytes(self) -> int:
        return {0: 1, 1: 1, 2: 1, 3: 2, 4: 2, 7: 2, 8: 2, 9: 3, 10: 4, 11: 4, 12: 6, 13: 8, 14: 4, 15: 4, 18: 8, 19: 16}[int(self)]


class Block(enum.IntEnum):
    Pixel = 0
    Yuv420Planar = 1  # I420; an addition of this backend (the reference lowers Block::Pixel only)
    Yuv420Nv12 = 2


class Whitepoint(enum.IntEnum):
    A = 0
    B = 1
    C = 2
    D50 = 3
    D55 = 4
    D65 = 5
    D75 = 6
    E = 7
    F2 = 8
    F7 = 9
    F11 = 10


class Primaries(enum.IntEnum):
    Bt709 = 0
    Bt601_525 = 1
    Bt601_625 = 2
    Smpte240 = 3
    Bt2020 = 4
    Bt2100 = 5


class ColorModel(enum.IntEnum):
    Rgb = 0
    Scalars = 1
    Oklab = 2
    SrLab2 = 3


class YuvMatrix(enum.IntEnum):
    Bt601 = 0
    Bt709 = 1
    Bt2020 = 2


class ColorChannel(enum.Enum):
    R = "R"
    G = "G"
    B = "B"
    Alpha = "A"
    Luma = "Luma"


@dataclass(frozen=True)
class Color:
    """image_canvas::color::Color (the variants the reference lowers)."""
    model: ColorModel
    transfer: Transfer = Transfer.Linear
    primary: Primaries = Primaries.Bt709
    whitepoint: Whitepoint = Whitepoint.D65

    @staticmethod
    def Rgb(primary: Primaries, transfer: Transfer, whitepoint: Whitepoint = Whitepoint.D65) -> "Color":
        return Color(ColorModel.Rgb, transfer, primary, whitepoint)

    @staticmethod
    def Scalars(transfer: Transfer = Transfer.Linear) -> "Color":
        return Color(ColorModel.Scalars, transfer)

    @staticmethod
    def SrLab2(whitepoint: Whitepoint) -> "Color":
        return Color(ColorModel.SrLab2, Transfer.Linear, Primaries.Bt709, whitepoint)


Color.SRGB = Color(ColorModel.Rgb, Transfer.Srgb, Primaries.Bt709, Whitepoint.D65)
Color.BT709_RGB = Color(ColorModel.Rgb, Transfer.Bt709, Primaries.Bt709, Whitepoint.D65)
Color.Oklab = Color(ColorModel.Oklab)


@dataclass(frozen=True)
class Texel:
    block: Block
    bits: SampleBits
    parts: SampleParts

    @staticmethod
    def new_u8(parts: SampleParts) -> "Texel":
        bits = {1: SampleBits.UInt8, 2: SampleBits.UInt8x2, 3: SampleBits.UInt8x3, 4: SampleBits.UInt8x4}[parts.num_components()]
        return Texel(Block.Pixel, bits, parts)

    @staticmethod
    def new_u16(parts: SampleParts) -> "Texel":
        bits = {1: SampleBits.UInt16, 2: SampleBits.UInt16x2, 3: SampleBits.UInt16x3, 4: SampleBits.UInt16x4}[parts.num_components()]
        return Texel(Block.Pixel, bits, parts)

    @staticmethod
    def new_f16(parts: SampleParts = SampleParts.RgbA) -> "Texel":
        return Texel(Block.Pixel, SampleBits.Float16x4, parts)

    @staticmethod
    def new_f32(parts: SampleParts = SampleParts.RgbA) -> "Texel":
        return Texel(Block.Pixel, SampleBits.Float32x4, parts)

    def channel_texel(self, channel: ColorChannel) -> Optional["Texel"]:
        """TexelExt::channel_texel, buffer.rs:57-62 (SampleParts::with_channel + matching bit depth)."""
        parts = {ColorChannel.R: SampleParts.R, ColorChannel.G: SampleParts.G, ColorChannel.B: SampleParts.B,
                 ColorChannel.Alpha: SampleParts.A, ColorChannel.Luma: SampleParts.Luma}.get(channel)
        if parts is None:
            return None
        bits = {SampleBits.UInt8x4: SampleBits.UInt8, SampleBits.UInt8x3: SampleBits.UInt8, SampleBits.UInt8x2: SampleBits.UInt8,
                SampleBits.UInt16x4: SampleBits.UInt16, SampleBits.UInt16x3: SampleBits.UInt16,
                SampleBits.UInt16x2: SampleBits.UInt16}.get(self.bits)
        if bits is None:
            return None
        return Texel(self.block, bits, parts)


@dataclass(frozen=True)
class ByteLayout:
    """buffer.rs:13-19 (host layout: tight rows unless stated otherwise)."""
    width: int
    height: int
    row_stride: int
    texel_stride: int


@dataclass(frozen=True)
class Descriptor:
    layout: ByteLayout
    color: Color
    texel: Texel
    # planar YUV only (this backend's addition)
    yuv_matrix: YuvMatrix = YuvMatrix.Bt709
    yuv_full_range: bool = False
    chroma_filter: int = 0

    EMPTY: "Descriptor" = None  # set below

    @staticmethod
    def with_texel(texel: Texel, width: int, height: int) -> Optional["Descriptor"]:
        """buffer.rs:95-115: Scalars/Linear colour, tight rows."""
        b = texel.bits.bytes()
        if width <= 0 or height <= 0 or width * b >= 2 ** 32:
            return None
        return Descriptor(ByteLayout(width, height, b * width, b), Color.Scalars(Transfer.Linear), texel)

    @staticmethod
    def with_srgb_image(kind: str, width: int, height: int) -> "Descriptor":
        """Descriptor::with_srgb_image for the image crate's colour types used by the tests."""
        texel = {"rgba8": Texel.new_u8(SampleParts.RgbA), "rgb8": Texel.new_u8(SampleParts.Rgb),
                 "luma8": Texel.new_u8(SampleParts.Luma), "luma_a8": Texel.new_u8(SampleParts.LumaA),
                 "rgba16": Texel.new_u16(SampleParts.RgbA), "luma16": Texel.new_u16(SampleParts.Luma),
                 "luma_a16": Texel.new_u16(SampleParts.LumaA)}[kind]
        b = texel.bits.bytes()
        return Descriptor(ByteLayout(width, height, b * width, b), Color.SRGB, texel)

    def is_consistent(self) -> bool:
        return self.texel.bits.bytes() == self.layout.texel_stride  # buffer.rs:165-168

    def size(self):
        return (self.layout.width, self.layout.height)

    def with_color(self, color: Color) -> "Descriptor":
        return replace(self, color=color)

    def to_aligned(self) -> ByteLayout:
        """buffer.rs:121-134: the 256-byte row pitch of the device copy."""
        stride = (self.layout.texel_stride * self.layout.width + 255) // 256 * 256
        return ByteLayout(self.layout.width, self.layout.height, stride, self.texel.bits.bytes())

    def chroma(self):
        return (self.texel, self.color)

    def to_ffi(self, row_stride: Optional[int] = None) -> _ffi.ZosDesc:
        d = _ffi.ZosDesc()
        d.width, d.height = self.layout.width, self.layout.height
        d.row_stride = self.to_aligned().row_stride if row_stride is None else row_stride
        d.texel_stride = self.layout.texel_stride
        d.block = int(self.texel.block)
        d.bits, d.parts = int(self.texel.bits), int(self.texel.parts)
        d.color, d.transfer = int(self.color.model), int(self.color.transfer)
        d.primaries, d.whitepoint = int(self.color.primary), int(self.color.whitepoint)
        d.yuv_matrix, d.yuv_full_range, d.chroma_filter = int(self.yuv_matrix), int(self.yuv_full_range), int(self.chroma_filter)
        return d


def yuv420_descriptor(width: int, height: int, color: Color, matrix: YuvMatrix = YuvMatrix.Bt709, full_range: bool = False,
                      nv12: bool = False, chroma_filter: int = 0) -> Descriptor:
    """Planar 8-bit 4:2:0 frames (I420 or NV12).  `color` gives primaries / transfer of the R'G'B' the
    matrix produces.  Not expressible in the reference (program.rs:794-938 lowers Block::Pixel only)."""
    texel = Texel(Block.Yuv420Nv12 if nv12 else Block.Yuv420Planar, SampleBits.UInt8, SampleParts.Yuv)
    return Descriptor(ByteLayout(width, height, width, 1), color, texel, matrix, full_range, chroma_filter)
