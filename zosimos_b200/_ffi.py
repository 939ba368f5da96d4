"""ctypes view of include/zosimos_cuda.h (the C-ABI of libzosimos_cuda.so).

The library is the product; this module only declares its structs and loads it.  There is no
fallback of any kind: if the shared object is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZOS_CUDA_LIB: another build of the same library (A/B runs of two kernel versions on one box, profiles/ab.sh)
LIB_PATH = os.environ.get("ZOS_CUDA_LIB") or os.path.join(_HERE, "libzosimos_cuda.so")

ZOS_MAX_STEPS = 8

# status codes
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_TYPE, ERR_STATE, ERR_OOM = range(7)
STATUS_NAMES = ["OK", "INVALID", "UNSUPPORTED", "CUDA", "TYPE", "STATE", "OOM"]

# step kinds
STEP_MATRIX, STEP_OKLAB_ENC, STEP_OKLAB_DEC, STEP_SRLAB2_ENC, STEP_SRLAB2_DEC, STEP_REQUANT, STEP_INJECT, STEP_F16 = range(1, 9)
# compose
SAMPLE_NEAREST, SAMPLE_BILINEAR = 0, 1
MAP_RECT, MAP_AFFINE, MAP_GRID8, MAP_SCALE = 0, 1, 2, 3
RUN_EAGER, RUN_GRAPH = 0, 1
BLEND_OVERWRITE = -1
(BLEND_CLEAR, BLEND_SRC, BLEND_DST, BLEND_SRC_OVER, BLEND_DST_OVER, BLEND_SRC_IN, BLEND_DST_IN, BLEND_SRC_OUT,
 BLEND_DST_OUT, BLEND_SRC_ATOP, BLEND_DST_ATOP, BLEND_XOR) = range(12)
BLEND_INJECT = 12
# program ops
GEN_BILINEAR, GEN_SOLID, GEN_NORMAL2D, GEN_FRACTAL_NOISE = 0, 1, 2, 3
OP_INPUT, OP_OUTPUT, OP_PIXEL, OP_COMPOSE, OP_COPY, OP_GENERATE, OP_BOX3, OP_PALETTE, OP_BUFFER_INIT, OP_FROM_BUFFER, OP_DYNAMIC = range(1, 12)
FUSE_EXACT, FUSE_WIDE, FUSE_NONE = 0, 1, 2
BLOCK_PIXEL, BLOCK_YUV420_PLANAR, BLOCK_YUV420_NV12 = 0, 1, 2


class ZosDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("row_stride", C.c_uint64), ("texel_stride", C.c_uint32),
                ("block", C.c_uint32), ("bits", C.c_uint32), ("parts", C.c_uint32), ("color", C.c_uint32),
                ("transfer", C.c_uint32), ("primaries", C.c_uint32), ("whitepoint", C.c_uint32),
                ("yuv_matrix", C.c_uint32), ("yuv_full_range", C.c_uint32), ("chroma_filter", C.c_uint32),
                ("reserved", C.c_uint32)]


class ZosTexFmt(C.Structure):
    _fields_ = [("transfer", C.c_uint32), ("parts", C.c_uint32), ("bits", C.c_uint32), ("storage", C.c_uint32)]


class ZosImage(C.Structure):
    _fields_ = [("desc", ZosDesc), ("data", C.c_void_p), ("plane1", C.c_void_p), ("plane2", C.c_void_p),
                ("chroma_stride", C.c_uint64), ("batch_stride", C.c_uint64), ("chroma_batch_stride", C.c_uint64)]


class ZosStep(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("fmt", ZosTexFmt), ("m", C.c_float * 9), ("v", C.c_float * 4)]


class ZosComposeParams(C.Structure):
    _fields_ = [("map", C.c_int32), ("sampling", C.c_int32), ("blend", C.c_int32), ("use_tma", C.c_int32),
                ("sel", C.c_int32 * 4), ("tgt", C.c_int32 * 4), ("inv", C.c_float * 9),
                ("dst_origin", C.c_int32 * 2), ("src_origin", C.c_int32 * 2), ("src_full", C.c_int32 * 2),
                ("inject_mix", C.c_float * 4), ("inject_color", C.c_float * 4),
                ("n_src_steps", C.c_uint32), ("n_dst_steps", C.c_uint32),
                ("src_steps", ZosStep * ZOS_MAX_STEPS), ("dst_steps", ZosStep * ZOS_MAX_STEPS)]


class ZosOp(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("src", C.c_int32 * 2), ("dst", C.c_int32), ("desc", ZosDesc),
                ("nsteps", C.c_uint32), ("steps", ZosStep * ZOS_MAX_STEPS), ("compose", ZosComposeParams),
                ("gen", C.c_float * 24), ("knob", C.c_uint32), ("reg", C.c_int32), ("data", C.c_void_p), ("data_len", C.c_uint64), ("source", C.c_char_p)]


class ZosArenaStats(C.Structure):
    _fields_ = [("device_allocs", C.c_uint64), ("reuses", C.c_uint64), ("bytes_reserved", C.c_uint64),
                ("bytes_in_use", C.c_uint64), ("bytes_parked", C.c_uint64)]


class ZosProgramStats(C.Structure):
    _fields_ = [("kernels", C.c_uint32), ("temp_buffers", C.c_uint32), ("temp_bytes", C.c_uint64), ("released", C.c_uint32),
                ("reserved", C.c_uint32), ("runs", C.c_uint64), ("graph_launches", C.c_uint64)]


class ZosError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("zosimos_cuda: %s: %s" % (STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else status, message))
        self.status = status
        self.message = message


# every symbol include/zosimos_cuda.h declares: (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "zos_abi_version": (C.c_uint32, []),
    "zos_bits_bytes": (C.c_uint32, [C.c_uint32]),
    "zos_aligned_row_stride": (C.c_uint64, [C.c_uint32, C.c_uint32]),
    "zos_desc_texfmt": (C.c_int32, [C.POINTER(ZosDesc), C.POINTER(ZosTexFmt)]),
    "zos_desc_device_bytes": (C.c_uint64, [C.POINTER(ZosDesc)]),
    "zos_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(_P)]),
    "zos_ctx_destroy": (None, [_P]),
    "zos_last_error": (C.c_char_p, [_P]),
    "zos_ctx_device": (C.c_int32, [_P]),
    "zos_ctx_stream": (_P, [_P]),
    "zos_sync": (C.c_int32, [_P]),
    "zos_ctx_set_flags": (C.c_int32, [_P, C.c_uint32]),
    "zos_ctx_launch_count": (C.c_uint64, [_P]),
    "zos_buf_alloc": (C.c_int32, [_P, C.c_uint64, C.POINTER(_P)]),
    "zos_buf_free": (None, [_P, _P]),
    "zos_buf_ptr": (_P, [_P]),
    "zos_buf_size": (C.c_uint64, [_P]),
    "zos_host_alloc": (C.c_int32, [_P, C.c_uint64, C.POINTER(_P)]),
    "zos_host_free": (None, [_P, _P]),
    "zos_buf_upload": (C.c_int32, [_P, _P, C.c_uint64, C.c_uint64, _P, C.c_uint64, C.c_uint64, C.c_uint64]),
    "zos_buf_download": (C.c_int32, [_P, _P, C.c_uint64, C.c_uint64, _P, C.c_uint64, C.c_uint64, C.c_uint64]),
    "zos_buf_copy": (C.c_int32, [_P, _P, C.c_uint64, _P, C.c_uint64, C.c_uint64]),
    "zos_buf_fill": (C.c_int32, [_P, _P, C.c_uint64, C.c_uint64, C.c_uint8]),
    "zos_image_upload": (C.c_int32, [_P, C.POINTER(ZosImage), C.c_uint32, _P]),
    "zos_image_download": (C.c_int32, [_P, C.POINTER(ZosImage), C.c_uint32, _P]),
    "zos_pixel_chain": (C.c_int32, [_P, C.POINTER(ZosImage), C.POINTER(ZosImage), C.POINTER(ZosStep), C.c_uint32, C.c_uint32]),
    "zos_compose": (C.c_int32, [_P, C.POINTER(ZosImage), C.POINTER(ZosImage), C.POINTER(ZosImage), C.POINTER(ZosComposeParams), C.c_uint32]),
    "zos_generate": (C.c_int32, [_P, C.POINTER(ZosImage), C.c_uint32, C.POINTER(C.c_float), C.c_uint32]),
    "zos_generate_from_buffer": (C.c_int32, [_P, C.POINTER(ZosImage), C.c_uint32, _P, C.c_uint64, C.c_uint32]),
    "zos_generate_bilinear": (C.c_int32, [_P, C.POINTER(ZosImage), C.POINTER(C.c_float), C.c_uint32]),
    "zos_generate_solid": (C.c_int32, [_P, C.POINTER(ZosImage), C.POINTER(C.c_float), C.c_uint32]),
    "zos_box3": (C.c_int32, [_P, C.POINTER(ZosImage), C.POINTER(ZosImage), C.POINTER(C.c_float), C.c_uint32]),
    "zos_palette": (C.c_int32, [_P, C.POINTER(ZosImage), C.POINTER(ZosImage), C.POINTER(ZosImage), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_uint32]),
    "zos_program_create": (C.c_int32, [_P, C.POINTER(ZosOp), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "zos_program_destroy": (None, [_P]),
    "zos_program_bind": (C.c_int32, [_P, C.c_int32, C.POINTER(ZosImage)]),
    "zos_program_unbind": (C.c_int32, [_P, C.c_int32]),
    "zos_program_set_knob": (C.c_int32, [_P, C.c_uint32, _P, C.c_uint64]),
    "zos_poll": (C.c_int32, [_P, C.POINTER(C.c_int32)]),
    "zos_program_reset_knobs": (C.c_int32, [_P]),
    "zos_multi_launch": (C.c_int32, [C.POINTER(_P), C.c_uint32, C.c_uint32]),
    "zos_multi_sync": (C.c_int32, [C.POINTER(_P), C.c_uint32]),
    "zos_gather_peer": (C.c_int32, [_P, _P, C.POINTER(C.c_uint64), C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint32]),
    "zos_comm_unique_id": (C.c_int32, [C.POINTER(C.c_uint8)]),
    "zos_comm_create": (C.c_int32, [_P, C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "zos_comm_destroy": (None, [_P]),
    "zos_comm_nccl_version": (C.c_int32, []),
    "zos_gather_nccl": (C.c_int32, [_P, _P, C.c_uint64, _P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int32]),
    "zos_program_launch": (C.c_int32, [_P]),
    "zos_program_step": (C.c_int32, [_P, C.c_uint32, C.POINTER(C.c_int32)]),
    "zos_program_kernel_count": (C.c_uint32, [_P]),
    "zos_affine_box_width": (C.c_int32, [C.c_int32, C.c_float, C.c_float]),
    "zos_srgb_encoder_tables": (C.c_int32, [C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "zos_dynamic_create": (C.c_int32, [_P, C.c_char_p, C.POINTER(_P)]),
    "zos_dynamic_launch": (C.c_int32, [_P, _P, C.POINTER(ZosImage), C.POINTER(ZosImage), C.POINTER(ZosImage), C.c_void_p, C.c_uint64]),
    "zos_program_run": (C.c_int32, [_P, C.c_uint32]),
    "zos_program_graph_launches": (C.c_uint64, [_P]),
    "zos_program_register_image": (C.c_int32, [_P, C.c_int32, C.POINTER(ZosImage)]),
    "zos_ctx_arena_stats": (C.c_int32, [_P, C.POINTER(ZosArenaStats)]),
    "zos_ctx_arena_trim": (C.c_int32, [_P]),
    "zos_program_release_buffers": (C.c_int32, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "zos_program_recover_buffers": (C.c_int32, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "zos_program_resources": (C.c_int32, [_P, C.POINTER(ZosProgramStats)]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads libzosimos_cuda.so (built in tree by __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the library lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib
