"""zosimos_b200 -- B200-native execution backend for zosimos compositing programs.

The product is libzosimos_cuda.so (hand-written sm_100a CUDA kernels behind the C-ABI of
include/zosimos_cuda.h).  This package is the host-side mirror of the reference's Rust API for the
hot path (CommandBuffer / Linker / Program / Executable / Execution / Pool, image-canvas
descriptors).  There is no CPU fallback: without the library or without a CUDA device, calls raise.
"""
from . import _ffi
from .buffer import (Block, ByteLayout, Color, ColorChannel, ColorModel, Descriptor, Primaries, SampleBits, SampleParts,
                     Texel, Transfer, Whitepoint, YuvMatrix, yuv420_descriptor)
from .device import Context, DeviceBuffer, DeviceImage, PinnedArray

__all__ = ["Block", "ByteLayout", "Color", "ColorChannel", "ColorModel", "Descriptor", "Primaries", "SampleBits",
           "SampleParts", "Texel", "Transfer", "Whitepoint", "YuvMatrix", "yuv420_descriptor", "Context",
           "DeviceBuffer", "DeviceImage", "PinnedArray"]
