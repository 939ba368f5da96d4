"""Shared helpers of the GPU parity tests: product descriptors <-> oracle descriptors."""
import numpy as np
import pytest

from oracle import oracle as O

import zosimos_b200 as Z
from zosimos_b200.buffer import ByteLayout, Color, ColorModel, Descriptor, Texel

B = Z.SampleBits


@pytest.fixture(scope="module")
def ctx():
    c = Z.Context(0)
    yield c
    c.close()


def zdesc(w, h, texel: Texel, color: Color) -> Descriptor:
    b = texel.bits.bytes()
    return Descriptor(ByteLayout(w, h, w * b, b), color, texel)


_PRIM = {Z.Primaries.Bt709: "bt709", Z.Primaries.Bt601_525: "bt601_525", Z.Primaries.Bt601_625: "bt601_625",
         Z.Primaries.Smpte240: "smpte240", Z.Primaries.Bt2020: "bt2020", Z.Primaries.Bt2100: "bt2100"}
_MODEL = {ColorModel.Rgb: "rgb", ColorModel.Scalars: "scalars", ColorModel.Oklab: "oklab", ColorModel.SrLab2: "srlab2"}


def to_oracle_color(c: Color) -> O.Color:
    return O.Color(_MODEL[c.model], int(c.transfer), _PRIM[c.primary], c.whitepoint.name)


def oracle_desc(d: Descriptor) -> O.Desc:
    return O.Desc(d.layout.width, d.layout.height, O.Texel(int(d.texel.bits), int(d.texel.parts)), to_oracle_color(d.color))


def oracle_image(d: Descriptor, data) -> O.Image:
    return O.Image(oracle_desc(d), np.ascontiguousarray(data).view(np.uint8).reshape(d.layout.height, -1))


def rand_bytes(h, row_bytes, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, row_bytes), dtype=np.uint8)


def gpu_image(ctx, desc, data):
    return ctx.upload(desc, data)
