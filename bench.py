#!/usr/bin/env python
"""bench.py -- throughput of the compositing hot path on B200 (contract: see the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|NAME[,NAME..]] [--impl reference]

The headline (top level of the JSON line) is BASELINE.json configs[1]: Porter-Duff source-over of two 3840x2160 RGBA8
sRGB layers composited in linear light.  With the default `--workload all` every other BASELINE configuration is timed
in the same invocation and reported under "workloads" (same keys: value, ms_per_step, roofline, clocks).

One "step" = LAUNCHES_PER_STEP launches of the workload's kernel, each over a batch of FRAMES frames resident in HBM.
The frames of one launch are distinct buffers far larger than the 126 MB L2 (c2_blend: 16 x 99.5 MB = 1.6 GB per launch),
so nothing is served from cache between launches.  LAUNCHES_PER_STEP is calibrated after the warm-up so that the K timed
steps last at least --min-seconds (default 1.0 s): the number is a sustained one, and the clock sampler has hundreds of
samples inside the region.  Frames are independent: with N GPUs every rank owns its own batch (weak scaling, no
collective on the data path).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline     algorithmic bytes per launch / CUDA-event kernel time vs the measured HBM peak; .burst = the same over the ten
               launches of the calibration (full SM clock), .traffic / .limiter = DRAM bytes and busiest unit from the committed
               ncu capture of this command (profiles/kernel_facts.json)
  gather       (N > 1) the optional collective of the path, timed alone: zos_gather_nccl of one output frame per rank
  cpu_baseline the CPU oracle (restatement of the reference pipeline; kind "port") on host cores
  e2e          the same metric through host buffers: pinned H2D of the layers + kernel + D2H (C-ABI calls)
  e2e_program  the same through the reference-shaped Program API (Executable.launch -> step -> Retire.output)
  workloads    the other BASELINE configurations + the reference's own `tests/loop.rs` loop (fps)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NOMINAL_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
HEADLINE = "c2_blend"
ALL_WORKLOADS = ["c2_blend", "c2_inscribe", "c1_oklab", "c3_affine_nearest", "c3_affine_bilinear", "c4_fused", "c5_rgba8", "c5_rgba16f",
                 "c5_rgb10a2", "c5_yuv420_yuv420", "c5_yuv420_rgba8", "loop_rs"]
DEFAULT_FRAMES = {"c2_blend": 16, "c2_inscribe": 16, "c1_oklab": 8, "c3_affine_nearest": 4, "c3_affine_bilinear": 4, "c4_fused": 64,
                  "c5_rgba8": 8, "c5_rgba16f": 8, "c5_rgb10a2": 8, "c5_yuv420_yuv420": 16, "c5_yuv420_rgba8": 16, "loop_rs": 1, "band_affine": 1}
C3_ANGLE_DEG = 30.0
C5_SIZE = (4096, 4096)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return NOMINAL_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def kernel_facts(name):
    """Per-kernel facts that only a profiler sees, from the committed ncu captures (profiles/kernel_facts.json, written by
    profiles/summarize.py from `ncu --set full` runs of this very command): DRAM traffic per launch and, for the kernels that
    are not HBM-bound, the instruction / SFU ceiling.  Never measured inside a timed run."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "kernel_facts.json"))).get(name) or {}
    except Exception:
        return {}


# ------------------------------------------------------------------ workload definitions (shared by both arms)
def workload_config(name):
    """The arm-independent description of a workload: the `config` object of the JSON line.  The GPU arm and the
    reference arm print the SAME dict for the same --workload."""
    W5, H5 = C5_SIZE
    cfg = {
        "c2_blend": dict(what="Porter-Duff source-over of two RGBA8 sRGB layers, composited in linear light", width=3840, height=2160,
                         bytes_per_px=12, baseline_config=1),
        "c2_inscribe": dict(what="inscribe of a full-size RGBA8 sRGB layer (overwrite incl. alpha)", width=3840, height=2160, bytes_per_px=8, baseline_config=1),
        "c1_oklab": dict(what="sRGB8 -> linear -> Oklab (LchA u8 register) -> sRGB8, one fused kernel", width=4096, height=4096, bytes_per_px=8, baseline_config=0),
        "c3_affine_nearest": dict(what="affine rotate %g deg about the centre, nearest, RGBA16F above over RGBA16F below" % C3_ANGLE_DEG, width=7680, height=4320,
                                  bytes_per_px=16, baseline_config=2),
        "c3_affine_bilinear": dict(what="affine rotate %g deg about the centre, bilinear, RGBA16F above over RGBA16F below (TMA-staged source tiles)" % C3_ANGLE_DEG,
                                   width=7680, height=4320, bytes_per_px=16, baseline_config=2),
        "c4_fused": dict(what="I420 1080p unpack -> BT.2020->709 matrix -> bilinear resize to 720p -> source-over on an RGBA8 sRGB canvas -> sRGB8 pack, one kernel",
                         width=1920, height=1080, out_width=1280, out_height=720, baseline_config=3,
                         bytes_per_frame="3.11 MB YUV in + 3.69 MB RGBA8 out = 6.80 MB (frames are opaque and cover the canvas, so the background is never read "
                                         "and is not counted)"),
        "c5_rgba8": dict(what="decode -> 3x3 primaries matrix -> encode, RGBA8 sRGB", width=W5, height=H5, bytes_per_px=8, baseline_config=4),
        "c5_rgba16f": dict(what="decode -> 3x3 primaries matrix -> encode, RGBA16F linear", width=W5, height=H5, bytes_per_px=16, baseline_config=4),
        "c5_rgb10a2": dict(what="decode -> 3x3 primaries matrix -> encode, RGB10A2 sRGB transfer (staged texel)", width=W5, height=H5, bytes_per_px=8, baseline_config=4),
        "c5_yuv420_yuv420": dict(what="I420 BT.2020 -> linear -> 3x3 -> I420 BT.709, one kernel", width=W5, height=H5, bytes_per_px=3, baseline_config=4),
        "c5_yuv420_rgba8": dict(what="I420 BT.2020 -> linear -> 3x3 -> RGBA8 sRGB, one kernel", width=W5, height=H5, bytes_per_px=5.5, baseline_config=4),
        "band_affine": dict(what="ONE 8192x8192 RGBA16F image, affine rotate 17 deg + scale, bilinear, split into N row bands (one per GPU), each band resampled "
                                 "from the source rows it needs; with and without a gather of the bands (zos_gather_nccl)", width=8192, height=8192, bytes_per_px=16,
                            baseline_config=4),
        "loop_rs": dict(what="the reference's tests/loop.rs: inscribe 157x151 on 512x512 RGBA8 sRGB, one pre-lowered Executable relaunched with host upload + "
                             "read-back every iteration (Program API)", width=512, height=512, baseline_config=None),
    }.get(name)
    if cfg is None:
        raise SystemExit("unknown workload " + name)
    out = {"workload": name}
    out.update(cfg)
    out["l2"] = "inputs larger than L2: the frames of one launch are distinct buffers (>= 435 MB per launch vs 126 MB L2), touched once"
    out["parity"] = PARITY_NOTE.get(name, PARITY_NOTE["default"])
    return out


PARITY_NOTE = {  # say next to every number what the result is pinned to
    "default": "semantics defined by this repo (the reference has no implementation); GPU == CPU oracle, oracle checked by property tests only",
    "c2_inscribe": "reference semantics; oracle reproduces the reference's `composed` golden hash (blockhash256), GPU == oracle bit-exact",
    "c1_oklab": "reference semantics; oracle reproduces the reference's `oklab` golden hash (blockhash256), GPU within 1 LSB of the oracle",
    "c3_affine_nearest": "reference semantics for nearest; oracle reproduces the `affine` golden hash on RGBA8; RGBA16F texel is ours; GPU == oracle bit-exact",
    "c5_rgba8": "reference semantics; colour matrices pinned by the `adapted` golden hash; GPU == oracle bit-exact",
    "c5_rgb10a2": "reference bit layout (stage.frag demux/mux, truncating pack); no reference test touches RGB10A2; GPU == oracle bit-exact",
    "loop_rs": "reference semantics; the GPU result hits the reference's `composed` golden hash and equals the oracle bit for bit",
}


class Workload:
    def __init__(self, name, out_px, in_px, bytes_per_frame, frames):
        self.name, self.out_px, self.in_px, self.bytes_per_frame, self.frames = name, out_px, in_px, bytes_per_frame, frames


def _descs():
    import zosimos_b200 as Z
    from zosimos_b200.buffer import ByteLayout, Color, Descriptor, SampleParts, Texel, Transfer

    def d(w, h, texel, color):
        b = texel.bits.bytes()
        return Descriptor(ByteLayout(w, h, w * b, b), color, texel)
    return Z, d, Color, Texel, SampleParts, Transfer


class HostParams:
    """The matrices the workloads need, from the product's host layer (zosh_to_xyz_matrix and Affine of
    libzosimos_cuda.so) and float64 numpy -- the GPU arm must not execute anything under oracle/."""

    @staticmethod
    def to_xyz(primaries, whitepoint):
        import zosimos_b200 as Z
        from zosimos_b200 import command
        prim = {"bt709": Z.Primaries.Bt709, "bt2020": Z.Primaries.Bt2020}[primaries]
        return command.to_xyz_matrix(prim, {"D65": Z.Whitepoint.D65}[whitepoint]).astype(np.float64)

    @staticmethod
    def inv3(m):
        return np.linalg.inv(np.asarray(m, dtype=np.float64).reshape(3, 3))

    @staticmethod
    def mul3(a, b):
        return np.asarray(a, dtype=np.float64).reshape(3, 3) @ np.asarray(b, dtype=np.float64).reshape(3, 3)

    @staticmethod
    def rotation_about(cx, cy, rad):
        """above -> below matrix of a rotation about (cx, cy): Affine::{shift, rotate, shift}, each a left multiplication."""
        from zosimos_b200.command import Affine, AffineSample
        a = Affine.new(AffineSample.Nearest).shift(-cx, -cy).rotate(rad).shift(cx, cy)
        return np.asarray(a.transformation, dtype=np.float32).reshape(3, 3)


def _replicate(ctx, img, one_frame):
    """Uploads one frame and copies it device-side into every frame slot of `img` (content does not influence the timing;
    generating gigabytes of host randomness would only slow the setup)."""
    tmp = ctx.image(img.desc, 1)
    tmp.upload(one_frame)
    for f in range(img.batch):
        ctx.check(ctx._lib.zos_buf_copy(ctx.handle, img.buf.handle, f * img.frame_bytes, tmp.buf.handle, 0, img.frame_bytes))
    ctx.sync()
    tmp.free()


def make_gpu_workload(name, ctx, frames, seed):
    """Returns (workload, launch(), images to free, extra).  Inputs are generated on the host with the seeds of
    SURVEY.md 8(d) and uploaded before timing."""
    Z, d, Color, Texel, SampleParts, Transfer = _descs()
    from zosimos_b200 import _ffi, ops
    O = HostParams()
    rng = np.random.default_rng(seed)
    rgba8 = Texel.new_u8(SampleParts.RgbA)
    lin = Color.Rgb(Z.Primaries.Bt709, Transfer.Linear)

    def rand_u8(fr, h, rb):
        return rng.integers(0, 256, (fr, h, rb), dtype=np.uint8)

    def yuv_source(W, H, matrix):
        sd = Z.yuv420_descriptor(W, H, Color.Rgb(Z.Primaries.Bt2020, Transfer.Bt709), matrix, False, False, 0)
        src = ctx.image(sd, frames)
        y = rng.integers(16, 236, (1, H, W), dtype=np.uint8)
        u = rng.integers(16, 241, (1, H // 2, W // 2), dtype=np.uint8); vv = rng.integers(16, 241, (1, H // 2, W // 2), dtype=np.uint8)
        one = ctx.image(sd, 1); one.upload((y, u, vv))
        for f in range(frames):
            ctx.check(ctx._lib.zos_buf_copy(ctx.handle, src.buf.handle, f * src.frame_bytes, one.buf.handle, 0, src.frame_bytes))
        ctx.sync(); one.free()
        return src

    if name in ("c2_blend", "c2_inscribe"):
        W, H = 3840, 2160
        desc = d(W, H, rgba8, Color.SRGB)
        below, above, dst = ctx.image(desc, frames), ctx.image(desc, frames), ctx.image(desc, frames)
        b0, a0 = rand_u8(1, H, W * 4), rand_u8(1, H, W * 4)
        _replicate(ctx, below, b0); _replicate(ctx, above, a0)
        blend = _ffi.BLEND_SRC_OVER if name == "c2_blend" else _ffi.BLEND_OVERWRITE
        p = ops.compose_params(blend=blend, sel=(0, 0, W, H), tgt=(0, 0, W, H))
        bpp = 12 if name == "c2_blend" else 8
        wl = Workload(name, W * H, W * H, W * H * bpp, frames)
        return wl, (lambda: ops.compose(ctx, below, above, dst, p)), [below, above, dst], dict(desc=desc, p=p, b0=b0, a0=a0, W=W, H=H)

    if name == "c1_oklab":
        W = H = 4096
        desc = d(W, H, rgba8, Color.SRGB)
        lch = d(W, H, Texel(Z.Block.Pixel, Z.SampleBits.UInt8x4, SampleParts.LchA), Color.Oklab)
        src, dst = ctx.image(desc, frames), ctx.image(desc, frames)
        _replicate(ctx, src, rand_u8(1, H, W * 4))
        T = O.to_xyz("bt709", "D65")
        steps = [ops.step(_ffi.STEP_OKLAB_ENC, T), ops.requant(lch), ops.step(_ffi.STEP_OKLAB_DEC, O.inv3(T))]
        return Workload(name, W * H, W * H, W * H * 8, frames), (lambda: ops.pixel_chain(ctx, src, dst, steps)), [src, dst], {}

    if name in ("c5_yuv420_yuv420", "c5_yuv420_rgba8"):
        W, H = C5_SIZE
        src = yuv_source(W, H, Z.YuvMatrix.Bt2020)
        M = O.mul3(O.inv3(O.to_xyz("bt709", "D65")), O.to_xyz("bt2020", "D65"))
        steps = [ops.matrix(M)]
        if name == "c5_yuv420_yuv420":
            dd = Z.yuv420_descriptor(W, H, Color.Rgb(Z.Primaries.Bt709, Transfer.Bt709), Z.YuvMatrix.Bt709, False, False, 0)
            dst = ctx.image(dd, frames)
            return Workload(name, W * H, W * H, W * H * 3, frames), (lambda: ops.pixel_chain(ctx, src, dst, steps)), [src, dst], {}
        dst = ctx.image(d(W, H, rgba8, Color.SRGB), frames)
        return Workload(name, W * H, W * H, W * H * 11 // 2, frames), (lambda: ops.pixel_chain(ctx, src, dst, steps)), [src, dst], {}

    if name in ("c5_rgba8", "c5_rgba16f", "c5_rgb10a2"):
        fmt = name[3:]
        W, H = C5_SIZE
        M = O.mul3(O.inv3(O.to_xyz("bt709", "D65")), O.to_xyz("bt2020", "D65"))
        if fmt == "rgba8":
            sd = d(W, H, rgba8, Color.SRGB); bpp = 8
        elif fmt == "rgba16f":
            sd = d(W, H, Texel.new_f16(), lin); bpp = 16
        else:
            sd = d(W, H, Texel(Z.Block.Pixel, Z.SampleBits.UInt1010102, SampleParts.RgbA), Color.Rgb(Z.Primaries.Bt709, Transfer.Srgb)); bpp = 8
        src, dst = ctx.image(sd, frames), ctx.image(sd, frames)
        data = rand_u8(1, H, W * sd.layout.texel_stride)
        if fmt == "rgba16f":
            data = rng.random((1, H, W * 4), dtype=np.float32).astype(np.float16).view(np.uint8)
        _replicate(ctx, src, data)
        steps = [ops.matrix(M)]
        return Workload(name, W * H, W * H, W * H * bpp, frames), (lambda: ops.pixel_chain(ctx, src, dst, steps)), [src, dst], {}

    if name in ("c3_affine_bilinear", "c3_affine_nearest"):
        W, H = 7680, 4320
        desc = d(W, H, Texel.new_f16(), lin)
        below, above, dst = ctx.image(desc, frames), ctx.image(desc, frames), ctx.image(desc, frames)
        v = rng.random((1, 270, W * 4), dtype=np.float32)
        v[rng.random(v.shape) < 0.01] *= 4.0
        tile = np.tile(v.astype(np.float16), (1, H // 270, 1)).view(np.uint8)
        _replicate(ctx, below, tile); _replicate(ctx, above, tile[:, ::-1].copy())
        m = O.rotation_about(W / 2, H / 2, float(np.deg2rad(C3_ANGLE_DEG)))
        inv = O.inv3(m.astype(np.float64)).astype(np.float32)
        p = ops.compose_params(map=_ffi.MAP_AFFINE, sampling=_ffi.SAMPLE_BILINEAR if name.endswith("bilinear") else _ffi.SAMPLE_NEAREST,
                               inv=inv, use_tma=True)
        return Workload(name, W * H, W * H, W * H * 16, frames), (lambda: ops.compose(ctx, below, above, dst, p)), [below, above, dst], {}

    if name == "c4_fused":
        W, H, w, h = 1920, 1080, 1280, 720
        src = yuv_source(W, H, Z.YuvMatrix.Bt709)
        od = d(w, h, rgba8, Color.SRGB)
        bg = ctx.upload(od, rand_u8(1, h, w * 4)[0])  # one canvas shared by all frames (batch_stride 0)
        dst = ctx.image(od, frames)
        M = O.mul3(O.inv3(O.to_xyz("bt709", "D65")), O.to_xyz("bt2020", "D65"))
        p = ops.compose_params(map=_ffi.MAP_SCALE, sampling=_ffi.SAMPLE_BILINEAR, blend=_ffi.BLEND_SRC_OVER, src_steps=[ops.matrix(M)], use_tma=True)
        # bytes that must move per frame: the YUV planes in, the RGBA8 frame out.  The canvas under an opaque, canvas-covering
        # frame is never read (k_frame_fast skips it), so it is NOT counted.
        bpf = W * H * 3 // 2 + w * h * 4
        return Workload(name, w * h, W * H, bpf, frames), (lambda: ops.compose(ctx, bg, src, dst, p)), [src, bg, dst], {}
    raise SystemExit("unknown workload " + name)


# ------------------------------------------------------------------ CPU side (oracle = "port"): cpu_baseline leg and --impl reference
def cpu_step_fn(name):
    """One bounded CPU sample of a workload through the oracle (the pass-structured restatement of the reference pipeline,
    oracle/zos_oracle.c with OpenMP over rows).  Returns (fn, output pixels per call, description of the sample)."""
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    M = O.mul3(O.inv3(O.to_xyz("bt709", "D65")), O.to_xyz("bt2020", "D65"))
    M32 = np.array(M, np.float32).reshape(3, 3)
    if name in ("c2_blend", "c2_inscribe"):
        W, H = 3840, 2160
        od = O.srgb_rgba8(W, H)
        a = O.Image(od, rng.integers(0, 256, (H, W * 4), dtype=np.uint8)); b = O.Image(od, rng.integers(0, 256, (H, W * 4), dtype=np.uint8))
        fn = (lambda: O.blend(b, (0, 0, W, H), a, 3)) if name == "c2_blend" else (lambda: O.inscribe(b, (0, 0, W, H), a, exact_quirks=False))
        return fn, W * H, "1 frame pair 3840x2160 per call"
    if name == "c1_oklab":
        W = H = 2048
        a = O.Image(O.srgb_rgba8(W, H), rng.integers(0, 256, (H, W * 4), dtype=np.uint8))
        return (lambda: O.color_convert(O.color_convert(a, O.OKLAB, O.Texel(O.B_UINT8X4, O.P_LCHA)), O.SRGB, O.RGBA8)), W * H, "2048x2048 (of 4096x4096) per call"
    if name in ("c3_affine_nearest", "c3_affine_bilinear"):
        W, H = 1920, 1080
        lin = O.Color("rgb", O.TR_LINEAR, "bt709", "D65")
        dsc = O.Desc(W, H, O.Texel(O.B_FLOAT16X4, O.P_RGBA), lin)
        data = rng.random((H, W * 4), dtype=np.float32).astype(np.float16).view(np.uint8)
        a = O.Image(dsc, data); b = O.Image(dsc, data[::-1].copy())
        m = O.shift(W / 2, H / 2) @ O.rotate(np.deg2rad(C3_ANGLE_DEG)) @ O.shift(-W / 2, -H / 2)
        return (lambda: O.affine(b, m, a, 1 if name.endswith("bilinear") else 0)), W * H, "1920x1080 RGBA16F (of 7680x4320) per call"
    if name == "c4_fused":
        W, H, w, h = 1920, 1080, 1280, 720
        y = rng.integers(16, 236, (H, W), dtype=np.uint8)
        u = rng.integers(16, 241, (H // 2, W // 2), dtype=np.uint8); v = rng.integers(16, 241, (H // 2, W // 2), dtype=np.uint8)
        od = O.srgb_rgba8(w, h)
        bg = O.decode(O.Image(od, rng.integers(0, 256, (h, w * 4), dtype=np.uint8)))

        def fn():
            tex = O.resize_pass(O.linear(O.decode_yuv420(y, u, v, W, H, 0.2126, 0.0722, False, False, 0, O.TR_BT709), M32), w, h, 1)
            canvas = bg.copy()
            O.blend_pass(canvas, tex, 0, 0, 3)
            return O.encode(od, canvas)
        return fn, w * h, "1 frame 1080p -> 720p per call"
    if name in ("c5_rgba8", "c5_rgba16f", "c5_rgb10a2"):
        W = H = 2048
        if name == "c5_rgba8":
            dsc = O.srgb_rgba8(W, H); data = rng.integers(0, 256, (H, W * 4), dtype=np.uint8)
        elif name == "c5_rgba16f":
            dsc = O.Desc(W, H, O.Texel(O.B_FLOAT16X4, O.P_RGBA), O.Color("rgb", O.TR_LINEAR, "bt709", "D65"))
            data = rng.random((H, W * 4), dtype=np.float32).astype(np.float16).view(np.uint8)
        else:
            dsc = O.Desc(W, H, O.Texel(O.B_UINT1010102, O.P_RGBA), O.Color("rgb", O.TR_SRGB, "bt709", "D65")); data = rng.integers(0, 256, (H, W * 4), dtype=np.uint8)
        a = O.Image(dsc, data)
        return (lambda: O.encode(dsc, O.linear(O.decode(a), M32))), W * H, "2048x2048 (of %dx%d) per call" % C5_SIZE
    if name in ("c5_yuv420_yuv420", "c5_yuv420_rgba8"):
        W = H = 2048
        y = rng.integers(16, 236, (H, W), dtype=np.uint8)
        u = rng.integers(16, 241, (H // 2, W // 2), dtype=np.uint8); v = rng.integers(16, 241, (H // 2, W // 2), dtype=np.uint8)
        od = O.srgb_rgba8(W, H)
        if name.endswith("rgba8"):
            return (lambda: O.encode(od, O.linear(O.decode_yuv420(y, u, v, W, H, 0.2627, 0.0593, False, False, 0, O.TR_BT709), M32))), W * H, "2048x2048 per call"
        return (lambda: O.encode_yuv420(O.linear(O.decode_yuv420(y, u, v, W, H, 0.2627, 0.0593, False, False, 0, O.TR_BT709), M32), 0.2126, 0.0722, False, O.TR_BT709)), \
            W * H, "2048x2048 per call"
    if name == "loop_rs":
        bgd = O.Image(O.srgb_rgba8(512, 512), rng.integers(0, 256, (512, 512 * 4), dtype=np.uint8))
        fgd = O.Image(O.srgb_rgba8(157, 151), rng.integers(0, 256, (151, 157 * 4), dtype=np.uint8))
        return (lambda: O.inscribe(bgd, (0, 0, 157, 151), fgd)), 512 * 512, "one 512x512 inscribe per call"
    raise SystemExit("unknown workload " + name)


def cpu_baseline(name, budget_s=12.0):
    """The CPU oracle timed in a FRESH interpreter: inside this process torch's OpenMP runtime is already
    loaded and the oracle's parallel regions end up on one thread (measured: 10x slower than the same
    code in a clean process), which would understate the CPU."""
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)  # torchrun sets it to 1
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-child", name, "--budget", str(budget_s)],
                             capture_output=True, text=True, timeout=budget_s * 6 + 120, env=env).stdout.strip().splitlines()
        return json.loads(out[-1])
    except Exception as e:  # never let the baseline break the bench line
        return {"value": None, "unit": "MP/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}


def cpu_baseline_child(name, budget_s=12.0):
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    fn, px, sample = cpu_step_fn(name)
    fn()  # warm up (page faults, tables)
    t0 = time.perf_counter(); n = 0
    while True:
        fn(); n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 200:
            break
    unit, v = ("fps", n / dt) if name == "loop_rs" else ("MP/s", px * n / dt / 1e6)
    return {"value": round(v, 2), "unit": unit, "cores": cores, "kind": "port",
            "sample": "%s, %d calls in %.1f s, oracle/zos_oracle.c with OpenMP on %d threads" % (sample, n, dt, cores)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Rust + wgpu
    (no cargo, no Vulkan ICD here), so oracle/_ref does not exist; the timed code is the oracle port."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = HEADLINE if args.workload == "all" else args.workload.split(",")[0]
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)  # all host threads (torchrun exports OMP_NUM_THREADS=1 to its workers); before libgomp loads
    fn, px, sample = cpu_step_fn(name)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    unit, v = ("fps", args.steps / dt) if name == "loop_rs" else ("MP/s", px * args.steps / dt / 1e6)
    v = round(v, 2)
    line = {"impl": "reference", "metric": "frames/sec" if name == "loop_rs" else "megapixels/sec", "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(name),
            "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port",
                             "sample": "each step = %s (a bounded sample of the batch the GPU arm processes per step); oracle/zos_oracle.c, OpenMP on %d threads" % (sample, cores)},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed regions.  NVML is polled directly (about 1 ms per
    sample); `nvidia-smi` (tens of ms per call) is the fallback."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.power, self.reasons, self.stop_flag, self.max_mhz, self.how = index, [], [], set(), False, None, None
        self.ready = threading.Event()  # set once the first poll is about to happen (NVML initialised), so that short regions get samples

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(self.index)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)  # survives CUDA_VISIBLE_DEVICES
        return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())

    def _run_nvml(self):
        nv, h = self._nvml_handle()
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
        self.how = "nvml"
        self.ready.set()
        while not self.stop_flag:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            try:
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            for nme, bit in zip(self.NAMES, bits):
                if r & bit:
                    self.reasons.add(nme)
            time.sleep(0.002)

    def _run_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        self.how = "nvidia-smi"
        self.ready.set()
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for nme, v in zip(self.NAMES, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            if not self.stop_flag:
                self._run_smi()

    def result(self):
        s = sorted(self.samples)
        out = {"sm_mhz": s[len(s) // 2] if s else None, "sm_min_mhz": s[0] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(s), "how": self.how}
        if self.power:
            out["power_w_max"] = round(max(self.power), 1)
        return out


class Sampling:
    """with Sampling(local, on) as s: ... ; s.result()"""

    def __init__(self, index, on):
        self.s = ClockSampler(index) if on else None

    def __enter__(self):
        if self.s:
            self.s.start()
            self.s.ready.wait(timeout=5)
        return self

    def __exit__(self, *a):
        if self.s:
            self.s.stop_flag = True
            self.s.join(timeout=2)

    def result(self):
        return self.s.result() if self.s else None


# ------------------------------------------------------------------ the GPU arm
class Rig:
    """What every timed region needs: the context, its stream as a torch stream, the barrier, max-over-ranks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        import zosimos_b200 as Z
        from zosimos_b200.shard import bind_to_gpu_numa
        self.Z = Z
        self.numa_cpus = bind_to_gpu_numa(self.local) if self.world > 1 else None  # staging memory local to each rank's GPU
        self.ctx = Z.Context(self.local)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.local)
        self.args = args

    def barrier(self, ctxs=()):
        for c in (self.ctx,) + tuple(ctxs):
            c.sync()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([float(v)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return type(v)(t.item())

    def event(self):
        return self.torch.cuda.Event(enable_timing=True)


def time_kernel_workload(rig, name, frames):
    """Warm-up, calibration of launches per step, K timed steps between CUDA events on the context's stream, max over ranks."""
    args, ctx = rig.args, rig.ctx
    wl, launch, images, extra = make_gpu_workload(name, ctx, frames, seed=1 + rank_seed(rig, name))
    warm = max(args.warmup, 3)
    for _ in range(warm):
        launch()
    rig.barrier()
    e0, e1 = rig.event(), rig.event()
    e0.record(rig.stream)
    for _ in range(10):  # calibration = the BURST figure: ten launches on a GPU at its full clock (a few milliseconds)
        launch()
    e1.record(rig.stream)
    rig.barrier()
    t_launch_ms = rig.max_over_ranks(e0.elapsed_time(e1) / 10.0)
    lps = max(1, int(math.ceil(args.min_seconds * 1e3 / (args.steps * t_launch_ms))))
    if args.launches_per_step:
        lps = args.launches_per_step
    with Sampling(rig.local, rig.rank == 0) as smp:
        l0 = ctx.launch_count
        ev = [rig.event() for _ in range(args.steps + 1)]
        ev[0].record(rig.stream)
        for i in range(args.steps):
            for _ in range(lps):
                launch()
            ev[i + 1].record(rig.stream)
        rig.barrier()
    launches = ctx.launch_count - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    total_ms_max = rig.max_over_ranks(total_ms)
    per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    peak, peak_src = hbm_peak()
    ms_step = total_ms_max / args.steps
    frames_all = rig.sum_over_ranks(wl.frames)  # (equal per rank unless --total-frames does not divide)
    value = wl.out_px * frames_all * lps / (ms_step * 1e-3) / 1e6
    k_ms = total_ms / (args.steps * lps)
    achieved = wl.bytes_per_frame * wl.frames / (k_ms * 1e-3) / 1e9
    facts = kernel_facts(name)
    roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
            "traffic": facts.get("traffic_bytes") if facts.get("frames") == wl.frames else None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": wl.bytes_per_frame * wl.frames, "kernel_ms_avg": round(k_ms, 4),
            "kernel_ms_median_step": round(per[len(per) // 2] / lps, 4)}
    b_ach = wl.bytes_per_frame * wl.frames / (t_launch_ms * 1e-3) / 1e9
    roof["burst"] = {"kernel_ms": round(t_launch_ms, 4), "achieved": round(b_ach, 1), "frac": round(b_ach / peak, 4),
                     "note": "10 launches back to back before the timed region (full SM clock; the timed region runs into the board's power cap)"}
    if facts.get("limiter"):  # the kernel's real ceiling when it is not HBM (instruction issue / SFU), from the ncu capture
        roof["limiter"] = facts["limiter"]
    res = {"value": round(value, 1), "unit": "MP/s", "ms_per_step": round(ms_step, 4), "launches_per_step": lps, "frames_per_launch_per_gpu": wl.frames,
           "timed_s": round(total_ms_max * 1e-3, 3), "gpu_launches": int(launches), "roofline": roof, "clocks": smp.result()}
    if wl.in_px != wl.out_px:
        res["value_input_px"] = round(wl.in_px * frames_all * lps / (ms_step * 1e-3) / 1e6, 1)
    return res, wl, images, extra


def rank_seed(rig, name):
    return rig.rank


def shard_frames(total, rank, world):
    from zosimos_b200.shard import frame_shard
    return frame_shard(total, rank, world)


def e2e_c_abi(rig, wl, extra):
    """End to end through the C-ABI with HOST buffers: per frame pinned host -> device (2 layers), kernel, device -> pinned
    host.  Three contexts (= three streams) on the device take frames round robin, so the upload of frame i+1, the kernel of
    frame i and the download of frame i-1 overlap on the copy engines."""
    import ctypes as C
    from zosimos_b200 import ops
    Z, ctx, torch = rig.Z, rig.ctx, rig.torch
    W, H, desc, p, b0, a0 = extra["W"], extra["H"], extra["desc"], extra["p"], extra["b0"], extra["a0"]
    fb = W * H * 4
    lib = ctx._lib
    lanes = []
    for _ in range(3):
        c = Z.Context(ctx.device)
        pin_in = c.pinned(2 * fb); pin_out = c.pinned(fb)
        pin_in.array[:fb] = b0.reshape(-1); pin_in.array[fb:] = a0.reshape(-1)
        lanes.append((c, pin_in, pin_out, c.image(desc, 1), c.image(desc, 1), c.image(desc, 1)))
    lane_ctxs = [l[0] for l in lanes]

    def frame(i):
        c, pin_in, pin_out, one_b, one_a, one_d = lanes[i % len(lanes)]
        c.check(lib.zos_buf_upload(c.handle, one_b.buf.handle, 0, one_b.pitch, C.c_void_p(pin_in.ptr.value), W * 4, W * 4, H))
        c.check(lib.zos_buf_upload(c.handle, one_a.buf.handle, 0, one_a.pitch, C.c_void_p(pin_in.ptr.value + fb), W * 4, W * 4, H))
        ops.compose(c, one_b, one_a, one_d, p)
        c.check(lib.zos_buf_download(c.handle, one_d.buf.handle, 0, one_d.pitch, C.c_void_p(pin_out.ptr.value), W * 4, W * 4, H))
    lane_streams = [torch.cuda.ExternalStream(c.stream, device=rig.local) for c in lane_ctxs]

    def region(nfr):
        t0 = rig.event()
        t1s = [rig.event() for _ in lane_streams]
        t0.record(lane_streams[0])
        for s in lane_streams[1:]:
            s.wait_event(t0)  # no lane starts before the start mark
        for i in range(nfr):
            frame(i)
        for s, e in zip(lane_streams, t1s):
            e.record(s)
        rig.barrier(lane_ctxs)
        return rig.max_over_ranks(max(t0.elapsed_time(e) for e in t1s))
    for i in range(6):
        frame(i)
    rig.barrier(lane_ctxs)
    t12 = region(12)
    nfr = max(12, int(math.ceil(rig.args.min_seconds * 1e3 / (t12 / 12.0) / 3.0)) * 3)
    with Sampling(rig.local, rig.rank == 0) as smp:
        ms = region(nfr)
    for l in lanes:
        for im in l[3:]:
            im.free()
        l[1].free(); l[2].free()
        l[0].close()
    val = wl.out_px * nfr * rig.world / (ms * 1e-3) / 1e6
    ceiling = None  # the raw pinned-copy ceiling of an 8-GPU box of this pool at this N (profiles/pcie_ceiling.py, no kernels), committed
    try:
        for l in open(os.path.join(ROOT, "profiles", "r02_pcie_ceiling_8gpu_box.jsonl")):
            r = json.loads(l)
            if r["n_gpus"] == rig.world:
                ceiling = {"raw_copy_ceiling": r["c2_blend_e2e_ceiling_mps"], "unit": "MP/s", "fraction": round(val / r["c2_blend_e2e_ceiling_mps"], 3),
                           "source": "profiles/r02_multi_gpu.md: pinned cudaMemcpyAsync both directions, %.1f GB/s aggregate at %d GPU(s) on that box" % (r["both_gbs_aggregate"], rig.world)}
    except Exception:
        pass
    return {"value": round(val, 1), "ceiling": ceiling, "unit": "MP/s", "h2d_bytes_per_step": 2 * fb * nfr, "d2h_bytes_per_step": fb * nfr, "frames": nfr, "timed_s": round(ms * 1e-3, 3),
            "pcie_gbs": {"h2d": round(2 * fb * nfr / ms / 1e6, 1), "d2h": round(fb * nfr / ms / 1e6, 1), "note": "per GPU, both directions concurrently"},
            "note": "per frame: zos_buf_upload x2 from pinned host, zos_compose, zos_buf_download to pinned host; 3 streams round robin so "
                    "copies overlap the kernels (PCIe bound)",
            "host_affinity": ("%d CPUs of the GPU's NUMA node" % len(rig.numa_cpus)) if rig.numa_cpus else "unbound", "clocks": smp.result()}


def gather_leg(rig, dst):
    """The optional collective of the path: every rank's FIRST output frame gathered to all ranks over NVLink with
    zos_gather_nccl (equal shards: ncclAllGather), timed alone with CUDA events on the context's stream, max over ranks."""
    from zosimos_b200 import shard
    ctx = rig.ctx
    comm = shard.Comm.from_torch_distributed(ctx)
    fb = dst.frame_bytes
    full = ctx.alloc(fb * rig.world)
    sizes, offs = [fb] * rig.world, [r * fb for r in range(rig.world)]
    for _ in range(3):
        comm.gather(dst.buf, 0, full, offs, sizes, -1)
    rig.barrier()
    n = 20
    e0, e1 = rig.event(), rig.event()
    e0.record(rig.stream)
    for _ in range(n):
        comm.gather(dst.buf, 0, full, offs, sizes, -1)
    e1.record(rig.stream)
    rig.barrier()
    ms = rig.max_over_ranks(e0.elapsed_time(e1)) / n
    comm.close()
    full.free()
    return {"what": "all-gather of one %.1f MB output frame per rank to every rank (zos_gather_nccl, NCCL %d)" % (fb / 1e6, ctx._lib.zos_comm_nccl_version()),
            "ms": round(ms, 4), "gbs_received_per_gpu": round(fb * (rig.world - 1) / ms / 1e6, 1), "in_timed_region": False}


def program_blend(rig, extra, pin):
    """c2_blend as a CommandBuffer program (input, input, blend, output) lowered once; returns run(n) -> ms for n relaunches
    through Executable.launch / Execution.step / Retire.output with HOST images of the pool (pageable numpy, or page-locked
    when `pin`)."""
    from zosimos_b200.command import Blend, CommandBuffer, Linker, Rectangle
    from zosimos_b200.program import Capabilities, Pool
    W, H, desc, b0, a0 = extra["W"], extra["H"], extra["desc"], extra["b0"], extra["a0"]
    pool = Pool(pin_host=pin)
    pool._devices.append(rig.ctx)  # the bench's context is the pool's device
    bg, fg = pool.insert(desc, b0), pool.insert(desc, a0)
    c = CommandBuffer()
    r0, r1 = c.input(desc), c.input(desc)
    res = c.blend(r0, Rectangle(0, 0, W, H), r1, Blend.Alpha)
    out, _ = c.output(res)
    exe = Linker.from_included().compile(c).lower_to(Capabilities.from_device(rig.ctx))
    keep = {"key": None}

    def once():
        env = exe.from_pool(pool)
        env.bind(r0, bg.key()); env.bind(r1, fg.key())
        if keep["key"] is not None:
            env.bind_output(out, keep["key"])  # the previous result image is written again (no host allocation per launch)
        ex = exe.launch(env)
        while ex.is_running():
            ex.step().block_on()
        ret = ex.retire_gracefully(pool)
        keep["key"] = ret.output(out).key()
        ret.finish()

    def run(n):
        t0 = time.perf_counter()
        for _ in range(n):
            once()
        rig.ctx.sync()
        return (time.perf_counter() - t0) * 1e3
    return run, exe, pool


def e2e_program(rig, wl, extra):
    """The product flow a user of the reference writes (tests/util.rs:85-118), timed by the host clock around whole
    launches: every launch uploads both layers from host images, runs the cached plan, reads the result back."""
    out = {}
    for pin in (False, True):
        run, exe, pool = program_blend(rig, extra, pin)
        run(3)
        rig.barrier()
        t3 = run(3)
        n = max(3, int(math.ceil(rig.args.min_seconds * 1e3 / (t3 / 3.0))))
        rig.barrier()
        ms = rig.max_over_ranks(run(n))
        out["pinned_host" if pin else "pageable_host"] = {"value": round(wl.out_px * n * rig.world / (ms * 1e-3) / 1e6, 1), "unit": "MP/s", "launches": n,
                                                          "ms_per_launch": round(ms / n, 3), "plans": exe.lowered}
        exe.close()
        pool.clear_cache()
    out["note"] = ("CommandBuffer(input, input, blend, output) lowered once; per launch: Executable.from_pool, bind x2, bind_output, launch, "
                   "step().block_on() until done, Retire.output (device -> host image), finish; serial, host clock")
    return out


def band_affine(rig):
    """SURVEY.md 8e, tile path: row bands of the OUTPUT of one very large image; every rank reads the source rows its band needs
    (the source is replicated), no halo exchange; the optional collective is the gather of the bands.  Strong scaling."""
    import ctypes as C
    from zosimos_b200 import _ffi, ops, shard
    from zosimos_b200.buffer import ByteLayout, Color, Descriptor, Texel, Transfer
    from zosimos_b200.command import Affine, AffineSample
    Z, ctx, args = rig.Z, rig.ctx, rig.args
    W = H = 8192
    lin = Color.Rgb(Z.Primaries.Bt709, Transfer.Linear)

    def desc(w, h):
        return Descriptor(ByteLayout(w, h, w * 8, 8), lin, Texel.new_f16())
    rng = np.random.default_rng(5)
    tile = rng.random((256, W * 4), dtype=np.float32).astype(np.float16)
    src = np.tile(tile, (H // 256, 1)).view(np.uint8)
    a = Affine.new(AffineSample.Nearest).shift(-W / 2, -H / 2).rotate(float(np.deg2rad(17.0))).scale(1.1, 0.9).shift(W / 2, H / 2)
    inv = np.linalg.inv(np.asarray(a.transformation, dtype=np.float64).reshape(3, 3)).astype(np.float32)
    bands = shard.row_bands(H, rig.world, 32)
    y0, y1 = bands[rig.rank]
    s0, s1 = shard.band_source_rows(inv.reshape(9), (y0, y1), W, H)
    above, below, dst = ctx.upload(desc(W, s1 - s0), src[s0:s1]), ctx.upload(desc(W, y1 - y0), src[y0:y1]), ctx.image(desc(W, y1 - y0))
    p = ops.compose_params(map=_ffi.MAP_AFFINE, sampling=_ffi.SAMPLE_BILINEAR, inv=inv, use_tma=True, dst_origin=(0, y0), src_origin=(0, s0), src_full=(W, H))
    comm = shard.Comm.from_torch_distributed(ctx) if rig.world > 1 else None
    full = ctx.alloc(H * W * 8) if comm else None
    sizes = [(b[1] - b[0]) * W * 8 for b in bands]
    offs = [b[0] * W * 8 for b in bands]

    def step(mode):
        ops.compose(ctx, below, above, dst, p)
        if comm and mode == "all":
            comm.gather(dst.buf, 0, full, offs, sizes, -1)
        elif comm and mode == "root":
            comm.gather(dst.buf, 0, full if rig.rank == 0 else None, offs if rig.rank == 0 else None, sizes, 0)
    out = {}
    for mode in (["off", "all", "root"] if comm else ["off"]):
        for _ in range(max(args.warmup, 3)):
            step(mode)
        rig.barrier()
        e0, e1 = rig.event(), rig.event()
        e0.record(rig.stream)
        for _ in range(3):
            step(mode)
        e1.record(rig.stream)
        rig.barrier()
        t = rig.max_over_ranks(e0.elapsed_time(e1) / 3.0)
        lps = max(1, int(math.ceil(args.min_seconds * 1e3 / (args.steps * t))))
        with Sampling(rig.local, rig.rank == 0) as smp:
            ev = [rig.event() for _ in range(args.steps + 1)]
            ev[0].record(rig.stream)
            for i in range(args.steps):
                for _ in range(lps):
                    step(mode)
                ev[i + 1].record(rig.stream)
            rig.barrier()
        ms = rig.max_over_ranks(ev[0].elapsed_time(ev[-1]))
        per_image_ms = ms / (args.steps * lps)
        out[mode] = {"value": round(W * H / (per_image_ms * 1e-3) / 1e6, 1), "unit": "MP/s", "ms_per_image": round(per_image_ms, 4), "launches_per_step": lps,
                     "timed_s": round(ms * 1e-3, 3), "clocks": smp.result()}
    peak, peak_src = hbm_peak()
    res = {"value": out["off"]["value"], "unit": "MP/s", "ms_per_step": round(out["off"]["ms_per_image"] * out["off"]["launches_per_step"], 4),
           "launches_per_step": out["off"]["launches_per_step"], "gpu_launches": args.steps * out["off"]["launches_per_step"],
           "roofline": {"bound": "hbm", "achieved": round(W * H * 16 / (out["off"]["ms_per_image"] * 1e-3) / 1e9, 1), "peak": peak * rig.world, "unit": "GB/s",
                        "frac": round(W * H * 16 / (out["off"]["ms_per_image"] * 1e-3) / 1e9 / (peak * rig.world), 4), "traffic": None, "peak_source": peak_src + " x n_gpus",
                        "algorithmic_bytes_per_launch": W * H * 16},
           "clocks": out["off"]["clocks"], "bands": bands,
           "gather": {"off": out["off"], "all_ranks": out.get("all"), "root_only": out.get("root"),
                      "note": "zos_gather_nccl on the context's stream after the band's kernel; image = %.0f MB, each rank contributes 1/N" % (W * H * 8 / 1e6)}}
    for im in (above, below, dst):
        im.free()
    if full:
        full.free()
    if comm:
        comm.close()
    return res


def loop_rs(rig):
    """The only number the reference publishes (tests/loop.rs:80-95, 'around 240 fps'): inscribe 157x151 on 512x512, one
    Executable, every iteration binds host images, launches, steps to the end and retires the output into the pool."""
    from zosimos_b200.buffer import Descriptor
    from zosimos_b200.command import CommandBuffer, Linker, Rectangle
    from zosimos_b200.program import Capabilities, Pool
    rng = np.random.default_rng(7)
    out = {}
    for pin in (False, True):
        pool = Pool(pin_host=pin)
        pool._devices.append(rig.ctx)
        bg = pool.insert_srgb(rng.integers(0, 256, (512, 512, 4), dtype=np.uint8))
        fg = pool.insert_srgb(rng.integers(0, 256, (151, 157, 4), dtype=np.uint8))
        c = CommandBuffer()
        r0, r1 = c.input(bg.descriptor()), c.input(fg.descriptor())
        res = c.inscribe(r0, Rectangle(0, 0, 157, 151), r1)
        o, _ = c.output(res)
        exe = Linker.from_included().compile(c).lower_to(Capabilities.from_device(rig.ctx))
        state = {"key": None}

        def once():
            env = exe.from_pool(pool)
            env.bind(r0, bg.key()); env.bind(r1, fg.key())
            if state["key"] is not None:
                env.bind_output(o, state["key"])
            ex = exe.launch(env)
            while ex.is_running():
                ex.step().block_on()
            ret = ex.retire_gracefully(pool)
            state["key"] = ret.output(o).key()
            ret.finish()
        for _ in range(20):
            once()
        rig.barrier()
        a0 = rig.ctx.arena_stats()["device_allocs"]
        l0 = rig.ctx.launch_count
        n = 200 * max(1, int(rig.args.min_seconds))
        with Sampling(rig.local, rig.rank == 0 and pin) as smp:
            t0 = time.perf_counter()
            for _ in range(n):
                once()
            rig.ctx.sync()
            dt = time.perf_counter() - t0
        dt = rig.max_over_ranks(dt)
        out["pinned_host" if pin else "pageable_host"] = {"fps": round(n * rig.world / dt, 1), "us_per_iteration": round(dt / n * 1e6, 1), "iterations": n,
                                                          "device_allocs_in_loop": rig.ctx.arena_stats()["device_allocs"] - a0,
                                                          "gpu_launches": rig.ctx.launch_count - l0, "plans": exe.lowered}
        if pin:
            out["clocks"] = smp.result()
        exe.close()
        pool.clear_cache()
    best = max(out["pinned_host"]["fps"], out["pageable_host"]["fps"])
    out.update({"value": best, "unit": "fps", "reference_published": "around 240 fps (tests/loop.rs:85, author's machine, 2021, hardware unstated)",
                "mp_per_s": round(best * 512 * 512 / 1e6, 1),
                "vs_baseline": round(best / rig.world / 240.0, 1),  # per GPU, against the one number BASELINE.md holds (other, unstated hardware)
                "note": "per iteration: Executable.from_pool, bind x2, bind_output, launch, step().block_on(), Retire.output (read-back into a pool host image), finish"})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="all", help="all | one name | comma list; the first one is the headline of the line")
    ap.add_argument("--frames", type=int, default=0, help="frames per launch per GPU (0 = the workload's default)")
    ap.add_argument("--total-frames", type=int, default=0, help="STRONG scaling: this many frames per launch over ALL GPUs (each rank takes total / N); "
                    "BASELINE config 3 is c4_fused with 256")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="each timed region lasts at least this long (launches per step are calibrated)")
    ap.add_argument("--launches-per-step", type=int, default=0, help="fix the launches per step instead of calibrating")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--angle", type=float, default=30.0, help="rotation of the c3_affine_* workloads in degrees (BASELINE: 30)")
    ap.add_argument("--size", default="4096x4096", help="image size WxH of the c5_* workloads (BASELINE config 5 sweeps 1-64 MP)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-baseline-child", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--budget", type=float, default=10.0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    global C3_ANGLE_DEG, C5_SIZE
    C3_ANGLE_DEG = args.angle
    C5_SIZE = tuple(int(v) for v in args.size.lower().split("x"))
    if args.cpu_baseline_child:
        print(json.dumps(cpu_baseline_child(args.cpu_baseline_child, args.budget)), flush=True)
        return
    if args.impl == "reference":
        return run_reference(args)

    names = ALL_WORKLOADS if args.workload == "all" else args.workload.split(",")
    for n in names:
        workload_config(n)  # unknown names fail before any GPU work
    rig = Rig(args)
    head = names[0]
    line = None
    others = {}
    total_launches = 0
    for name in names:
        frames = args.frames or DEFAULT_FRAMES[name]
        if args.total_frames and name not in ("loop_rs", "band_affine"):
            frames = len(shard_frames(args.total_frames, rig.rank, rig.world))
            if frames == 0:
                raise SystemExit("--total-frames %d leaves rank %d without work" % (args.total_frames, rig.rank))
        if name in ("loop_rs", "band_affine"):
            res = loop_rs(rig) if name == "loop_rs" else band_affine(rig)
            total_launches += res.get("gpu_launches", 0) if name == "band_affine" else 0
            res["config"] = workload_config(name)
        else:
            res, wl, images, extra = time_kernel_workload(rig, name, frames)
            total_launches += res["gpu_launches"]
            if name == head:
                if name == "c2_blend" and not args.no_e2e:
                    res["e2e"] = e2e_c_abi(rig, wl, extra)
                    res["e2e_program"] = e2e_program(rig, wl, extra)
                    res["gather"] = gather_leg(rig, images[2]) if rig.world > 1 else None
            for im in images:
                im.free()
            rig.ctx.arena_trim()
            res["config"] = workload_config(name)
        if name == head:
            line = res
        else:
            others[name] = res
    if rig.rank == 0:
        cfg = line.pop("config")
        top = {"metric": "frames/sec" if head == "loop_rs" else "megapixels/sec", "value": line.pop("value"), "unit": line.pop("unit"), "n_gpus": rig.world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": line.pop("ms_per_step", None), "higher_is_better": True,
               "scaling": "strong" if (args.total_frames or head == "band_affine") else "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg}
        top.update(line)
        top["parallelism"] = ("row bands of one image over %d GPU(s)" if head == "band_affine" else "frame-batch sharding over %d GPU(s), no collective on the data path") % rig.world
        if args.total_frames:
            top["total_frames_per_launch"] = args.total_frames
        if not args.no_cpu:
            top["cpu_baseline"] = cpu_baseline(head, args.budget)
            for n, r in others.items():  # a short sample each, so that every row has its CPU number beside it
                r["cpu_baseline"] = cpu_baseline(n, 1.5)
        if others:
            # the headline once more under its own name, so that `workloads` lists every configuration
            mine = {k: top[k] for k in ("value", "unit", "ms_per_step", "launches_per_step", "frames_per_launch_per_gpu", "timed_s", "gpu_launches", "roofline", "clocks",
                                        "config", "cpu_baseline") if k in top}
            top["workloads"] = {head: mine, **others}
            top["gpu_launches_all_workloads"] = total_launches
        print(json.dumps(top), flush=True)
    if rig.world > 1:
        rig.dist.barrier()
        rig.dist.destroy_process_group()
    rig.ctx.close()


if __name__ == "__main__":
    main()
