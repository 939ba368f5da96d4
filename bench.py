#!/usr/bin/env python
"""bench.py -- throughput of the compositing hot path on B200 (contract: see the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Default workload = BASELINE.json configs[1]: Porter-Duff source-over of two 3840x2160 RGBA8 sRGB
layers composited in linear light.  One "step" = one launch of the fused kernel over a batch of
FRAMES frame pairs resident in HBM (FRAMES * 99.5 MB touched once per step, far larger than the
126 MB L2, so nothing is served from cache between steps).  Frames are independent: with N GPUs
every rank owns its own batch (weak scaling, no collective on the data path).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline     algorithmic bytes per launch / CUDA-event kernel time vs the measured HBM peak
  cpu_baseline the CPU oracle (restatement of the reference pipeline; kind "port") on host cores
  e2e          the same metric through host buffers: pinned H2D of the layers + kernel + D2H
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NOMINAL_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return NOMINAL_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ workloads
class Workload:
    """name, output pixels per frame, algorithmic bytes per frame, builders for GPU and CPU."""

    def __init__(self, name, desc, out_px, in_px, bytes_per_frame, frames):
        self.name, self.desc, self.out_px, self.in_px, self.bytes_per_frame, self.frames = name, desc, out_px, in_px, bytes_per_frame, frames


def _descs():
    import zosimos_b200 as Z
    from zosimos_b200.buffer import ByteLayout, Color, Descriptor, SampleParts, Texel, Transfer

    def d(w, h, texel, color):
        b = texel.bits.bytes()
        return Descriptor(ByteLayout(w, h, w * b, b), color, texel)
    return Z, d, Color, Texel, SampleParts, Transfer


class HostParams:
    """The matrices the workloads need, from the product's host layer (zosh_to_xyz_matrix and Affine of
    libzosimos_cuda.so) and float64 numpy -- this arm must not execute anything under oracle/."""

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


def make_gpu_workload(name, ctx, frames, seed):
    """Returns (workload, launch(), e2e_step() or None).  Inputs are generated on the host with the
    seeds of SURVEY.md 8(d) and uploaded before timing."""
    Z, d, Color, Texel, SampleParts, Transfer = _descs()
    from zosimos_b200 import _ffi, ops
    O = HostParams()  # parameter blocks come from the product's own host layer; the oracle is not touched by this arm
    rng = np.random.default_rng(seed)
    rgba8 = Texel.new_u8(SampleParts.RgbA)
    lin = Color.Rgb(Z.Primaries.Bt709, Transfer.Linear)

    def rand_u8(fr, h, rb):
        return rng.integers(0, 256, (fr, h, rb), dtype=np.uint8)

    if name in ("c2_blend", "c2_inscribe"):
        W, H = 3840, 2160
        desc = d(W, H, rgba8, Color.SRGB)
        # one random frame pair, replicated on the device into `frames` distinct buffers (content does
        # not influence the timing; generating 1.6 GB of host randomness would only slow the setup)
        below, above, dst = ctx.image(desc, frames), ctx.image(desc, frames), ctx.image(desc, frames)
        b0, a0 = rand_u8(1, H, W * 4), rand_u8(1, H, W * 4)
        _replicate(ctx, below, b0); _replicate(ctx, above, a0)
        blend = _ffi.BLEND_SRC_OVER if name == "c2_blend" else _ffi.BLEND_OVERWRITE
        p = ops.compose_params(blend=blend, sel=(0, 0, W, H), tgt=(0, 0, W, H))
        bpp = 12 if name == "c2_blend" else 8
        wl = Workload(name, "Porter-Duff source-over, linear light" if name == "c2_blend" else "inscribe (full-size layer)",
                      W * H, W * H, W * H * bpp, frames)

        def launch():
            ops.compose(ctx, below, above, dst, p)

        # e2e: host layers -> device -> kernel -> host result, frame by frame through pinned memory
        # Three contexts (= three streams) on the same device take frames round robin, so the upload of
        # frame i+1, the kernel of frame i and the download of frame i-1 overlap on the copy engines.
        fb = W * H * 4
        import ctypes as C
        lib = ctx._lib
        lanes = []
        for _ in range(3):
            c = Z.Context(ctx.device)
            pin_in = c.pinned(2 * fb); pin_out = c.pinned(fb)
            pin_in.array[:fb] = b0.reshape(-1); pin_in.array[fb:] = a0.reshape(-1)
            lanes.append((c, pin_in, pin_out, c.image(desc, 1), c.image(desc, 1), c.image(desc, 1)))

        def e2e_frame(i):
            c, pin_in, pin_out, one_b, one_a, one_d = lanes[i % len(lanes)]
            c.check(lib.zos_buf_upload(c.handle, one_b.buf.handle, 0, one_b.pitch, C.c_void_p(pin_in.ptr.value), W * 4, W * 4, H))
            c.check(lib.zos_buf_upload(c.handle, one_a.buf.handle, 0, one_a.pitch, C.c_void_p(pin_in.ptr.value + fb), W * 4, W * 4, H))
            ops.compose(c, one_b, one_a, one_d, p)
            c.check(lib.zos_buf_download(c.handle, one_d.buf.handle, 0, one_d.pitch, C.c_void_p(pin_out.ptr.value), W * 4, W * 4, H))
        return wl, launch, (e2e_frame, 2 * fb, fb, [l[0] for l in lanes])

    if name == "c1_oklab":
        W = H = 4096
        desc = d(W, H, rgba8, Color.SRGB)
        lch = d(W, H, Texel(Z.Block.Pixel, Z.SampleBits.UInt8x4, SampleParts.LchA), Color.Oklab)
        src, dst = ctx.image(desc, frames), ctx.image(desc, frames)
        _replicate(ctx, src, rand_u8(1, H, W * 4))
        T = O.to_xyz("bt709", "D65")
        steps = [ops.step(_ffi.STEP_OKLAB_ENC, T), ops.requant(lch), ops.step(_ffi.STEP_OKLAB_DEC, O.inv3(T))]
        wl = Workload(name, "sRGB8 -> Oklab (LchA u8 register) -> sRGB8, one fused kernel", W * H, W * H, W * H * 8, frames)
        return wl, (lambda: ops.pixel_chain(ctx, src, dst, steps)), None

    if name in ("c5_yuv420_yuv420", "c5_yuv420_rgba8"):
        W, H = C5_SIZE
        sd = Z.yuv420_descriptor(W, H, Color.Rgb(Z.Primaries.Bt2020, Transfer.Bt709), Z.YuvMatrix.Bt2020, False, False, 0)
        src = ctx.image(sd, frames)
        y = rng.integers(16, 236, (1, H, W), dtype=np.uint8)
        u = rng.integers(16, 241, (1, H // 2, W // 2), dtype=np.uint8); vv = rng.integers(16, 241, (1, H // 2, W // 2), dtype=np.uint8)
        one = ctx.image(sd, 1); one.upload((y, u, vv))
        for f in range(frames):
            ctx.check(ctx._lib.zos_buf_copy(ctx.handle, src.buf.handle, f * src.frame_bytes, one.buf.handle, 0, src.frame_bytes))
        M = O.mul3(O.inv3(O.to_xyz("bt709", "D65")), O.to_xyz("bt2020", "D65"))
        if name == "c5_yuv420_yuv420":
            dd = Z.yuv420_descriptor(W, H, Color.Rgb(Z.Primaries.Bt709, Transfer.Bt709), Z.YuvMatrix.Bt709, False, False, 0)
            dst = ctx.image(dd, frames)
            wl = Workload(name, "I420 BT.2020 -> linear -> 3x3 -> I420 BT.709, one kernel", W * H, W * H, W * H * 3, frames)
            return wl, (lambda: ops.pixel_chain(ctx, src, dst, [ops.matrix(M)])), None
        dd = d(W, H, rgba8, Color.SRGB)
        dst = ctx.image(dd, frames)
        wl = Workload(name, "I420 BT.2020 -> linear -> 3x3 -> RGBA8 sRGB, one kernel", W * H, W * H, W * H * 11 // 2, frames)
        return wl, (lambda: ops.pixel_chain(ctx, src, dst, [ops.matrix(M)])), None

    if name.startswith("c5_"):
        fmt = name[3:]
        W, H = C5_SIZE
        M = O.mul3(O.inv3(O.to_xyz("bt709", "D65")), O.to_xyz("bt2020", "D65"))
        if fmt == "rgba8":
            sd = dd = d(W, H, rgba8, Color.SRGB); bpp = 8
        elif fmt == "rgba16f":
            sd = dd = d(W, H, Texel.new_f16(), lin); bpp = 16
        elif fmt == "rgb10a2":
            sd = dd = d(W, H, Texel(Z.Block.Pixel, Z.SampleBits.UInt1010102, SampleParts.RgbA), Color.Rgb(Z.Primaries.Bt709, Transfer.Srgb)); bpp = 8
        else:
            raise SystemExit("unknown workload " + name)
        src, dst = ctx.image(sd, frames), ctx.image(dd, frames)
        data = rand_u8(1, H, W * sd.layout.texel_stride)
        if fmt == "rgba16f":
            data = rng.random((1, H, W * 4), dtype=np.float32).astype(np.float16).view(np.uint8)
        _replicate(ctx, src, data)
        steps = [ops.matrix(M)]
        wl = Workload(name, "decode -> 3x3 primaries matrix -> encode (%s)" % fmt, W * H, W * H, W * H * bpp, frames)
        return wl, (lambda: ops.pixel_chain(ctx, src, dst, steps)), None

    if name in ("c3_affine_bilinear", "c3_affine_nearest"):
        W, H = 7680, 4320
        desc = d(W, H, Texel.new_f16(), lin)
        below, above, dst = ctx.image(desc, frames), ctx.image(desc, frames), ctx.image(desc, frames)
        v = rng.random((1, 270, W * 4), dtype=np.float32)
        v[rng.random(v.shape) < 0.01] *= 4.0
        tile = np.tile(v.astype(np.float16), (1, H // 270, 1)).view(np.uint8)
        _replicate(ctx, below, tile); _replicate(ctx, above, tile[:, ::-1].copy())
        ang = np.deg2rad(C3_ANGLE_DEG)
        m = O.rotation_about(W / 2, H / 2, float(ang))
        inv = O.inv3(m.astype(np.float64)).astype(np.float32)
        p = ops.compose_params(map=_ffi.MAP_AFFINE, sampling=_ffi.SAMPLE_BILINEAR if name.endswith("bilinear") else _ffi.SAMPLE_NEAREST,
                               inv=inv, use_tma=True)
        wl = Workload(name, "affine rotate %gdeg, %s, RGBA16F over RGBA16F" % (C3_ANGLE_DEG, name.split("_")[-1]), W * H, W * H, W * H * 16, frames)
        return wl, (lambda: ops.compose(ctx, below, above, dst, p)), None

    if name == "c4_fused":
        W, H, w, h = 1920, 1080, 1280, 720
        yuv = Z.yuv420_descriptor(W, H, Color.Rgb(Z.Primaries.Bt2020, Transfer.Bt709), Z.YuvMatrix.Bt709, False, False, 0)
        src = ctx.image(yuv, frames)
        y = rng.integers(16, 236, (1, H, W), dtype=np.uint8)
        u = rng.integers(16, 241, (1, H // 2, W // 2), dtype=np.uint8); vv = rng.integers(16, 241, (1, H // 2, W // 2), dtype=np.uint8)
        one = ctx.image(yuv, 1); one.upload((y, u, vv))
        for f in range(frames):
            ctx.check(ctx._lib.zos_buf_copy(ctx.handle, src.buf.handle, f * src.frame_bytes, one.buf.handle, 0, src.frame_bytes))
        od = d(w, h, rgba8, Color.SRGB)
        bg = ctx.upload(od, rand_u8(1, h, w * 4)[0])  # one background shared by all frames (batch_stride 0)
        dst = ctx.image(od, frames)
        M = O.mul3(O.inv3(O.to_xyz("bt709", "D65")), O.to_xyz("bt2020", "D65"))
        p = ops.compose_params(map=_ffi.MAP_SCALE, sampling=_ffi.SAMPLE_BILINEAR, blend=_ffi.BLEND_SRC_OVER, src_steps=[ops.matrix(M)], use_tma=True)
        bpf = W * H * 3 // 2 + 2 * w * h * 4
        wl = Workload(name, "I420 unpack -> BT.2020->709 matrix -> bilinear 1080p->720p -> over RGBA8 bg -> sRGB8 pack", w * h, W * H, bpf, frames)
        return wl, (lambda: ops.compose(ctx, bg, src, dst, p)), None
    raise SystemExit("unknown workload " + name)


def _replicate(ctx, img, one_frame):
    """Uploads one frame and copies it device-side into every frame slot of `img`."""
    import zosimos_b200 as Z
    tmp = ctx.image(img.desc, 1)
    tmp.upload(one_frame)
    for f in range(img.batch):
        ctx.check(ctx._lib.zos_buf_copy(ctx.handle, img.buf.handle, f * img.frame_bytes, tmp.buf.handle, 0, img.frame_bytes))
    ctx.sync()
    tmp.free()


# ------------------------------------------------------------------ CPU baseline (oracle = "port")
def cpu_baseline(name, budget_s=12.0):
    """The CPU oracle timed in a FRESH interpreter: inside this process torch's OpenMP runtime is already
    loaded and the oracle's parallel regions end up on one thread (measured: 10x slower than the same
    code in a clean process), which would understate the CPU."""
    if name not in ("c2_blend", "c2_inscribe", "c1_oklab"):
        return None
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)  # torchrun sets it to 1
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-child", name, "--budget", str(budget_s)],
                             capture_output=True, text=True, timeout=budget_s * 6 + 120, env=env).stdout.strip().splitlines()
        return json.loads(out[-1])
    except Exception as e:  # never let the baseline break the bench line
        return {"value": None, "unit": "MP/s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}


def cpu_baseline_child(name, budget_s=12.0):
    """Times the CPU oracle (pass-structured restatement of the reference pipeline, OpenMP over rows)
    on a bounded sample of the workload.  Returns MP/s (output pixels of the same definition)."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    rng = np.random.default_rng(1)
    if name in ("c2_blend", "c2_inscribe"):
        W, H = 3840, 2160
        od = O.srgb_rgba8(W, H)
        a = O.Image(od, rng.integers(0, 256, (H, W * 4), dtype=np.uint8)); b = O.Image(od, rng.integers(0, 256, (H, W * 4), dtype=np.uint8))
        fn = (lambda: O.blend(b, (0, 0, W, H), a, 3)) if name == "c2_blend" else (lambda: O.inscribe(b, (0, 0, W, H), a, exact_quirks=False))
        px = W * H; sample = "1 frame pair 3840x2160 per repetition"
    elif name == "c1_oklab":
        W = H = 2048
        od = O.srgb_rgba8(W, H)
        a = O.Image(od, rng.integers(0, 256, (H, W * 4), dtype=np.uint8))
        fn = lambda: O.color_convert(O.color_convert(a, O.OKLAB, O.Texel(O.B_UINT8X4, O.P_LCHA)), O.SRGB, O.RGBA8)
        px = W * H; sample = "2048x2048 per repetition"
    else:
        return None
    fn()  # warm up (page faults, tables)
    t0 = time.perf_counter(); n = 0
    while True:
        fn(); n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 50:
            break
    return {"value": round(px * n / dt / 1e6, 2), "unit": "MP/s", "cores": cores, "kind": "port",
            "sample": "%s, %d repetitions in %.1f s, oracle/zos_oracle.c with OpenMP on %d threads" % (sample, n, dt, cores)}


C3_ANGLE_DEG = 30.0
C5_SIZE = (4096, 4096)


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed regions.  NVML is polled directly (about 1 ms per
    sample, so even a few-millisecond region gets samples); `nvidia-smi` (tens of ms per call) is the fallback."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.how = index, [], set(), False, None, None
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
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            for nme, bit in zip(self.NAMES, bits):
                if r & bit:
                    self.reasons.add(nme)
            time.sleep(0.001)

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
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s),
                "how": self.how}


# ------------------------------------------------------------------ main
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Rust + wgpu
    (no cargo, no Vulkan ICD here), so oracle/_ref does not exist; the timed code is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)  # all host threads (torchrun exports OMP_NUM_THREADS=1 to its workers); before libgomp loads
    from oracle import oracle as O
    W, H = 3840, 2160
    rng = np.random.default_rng(1)
    od = O.srgb_rgba8(W, H)
    a = O.Image(od, rng.integers(0, 256, (H, W * 4), dtype=np.uint8)); b = O.Image(od, rng.integers(0, 256, (H, W * 4), dtype=np.uint8))
    if name == "c2_inscribe":
        fn = lambda: O.inscribe(b, (0, 0, W, H), a, exact_quirks=False)
    else:
        name = "c2_blend"
        fn = lambda: O.blend(b, (0, 0, W, H), a, 3)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = time.perf_counter() - t0
    v = round(W * H * args.steps / dt / 1e6, 2)
    sample = "each step = 1 frame pair 3840x2160 (of the %d-frame batch the GPU arm processes per step)" % args.frames
    line = {"impl": "reference", "metric": "megapixels/sec", "value": v, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "width": W, "height": H, "layers": 2, "frames_per_step": 1},
            "cpu_baseline": {"value": v, "unit": "MP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2_blend")
    ap.add_argument("--frames", type=int, default=16, help="frames per step per GPU")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--angle", type=float, default=30.0, help="rotation of the c3_affine_* workloads in degrees (BASELINE: 30)")
    ap.add_argument("--size", default="4096x4096", help="image size WxH of the c5_* workloads (BASELINE config 5 sweeps 1-64 MP)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-baseline-child", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--budget", type=float, default=12.0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_child:
        print(json.dumps(cpu_baseline_child(args.cpu_baseline_child, args.budget)), flush=True)
        return
    if args.impl == "reference":
        return run_reference(args)
    global C3_ANGLE_DEG, C5_SIZE
    C3_ANGLE_DEG = args.angle
    C5_SIZE = tuple(int(v) for v in args.size.lower().split("x"))

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import zosimos_b200 as Z
    from zosimos_b200.shard import bind_to_gpu_numa
    numa_cpus = bind_to_gpu_numa(local) if world > 1 else None  # staging memory local to each rank's GPU
    ctx = Z.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    wl, launch, e2e = make_gpu_workload(args.workload, ctx, args.frames, seed=1 + rank)

    def barrier():
        ctx.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ctx.sync(); torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        launch()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.ready.wait(timeout=5)
    l0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for i in range(args.steps):
        launch()
        ev[i + 1].record(stream)
    barrier()
    launches = ctx.launch_count - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())

    # end to end: pinned host -> device -> kernel -> host, every frame of the step
    e2e_out = None
    if e2e is not None:
        fn, h2d, d2h, lane_ctxs = e2e
        lane_streams = [torch.cuda.ExternalStream(c.stream, device=local) for c in lane_ctxs]

        def sync_lanes():
            for c in lane_ctxs:
                c.sync()
        for i in range(6):
            fn(i)
        sync_lanes(); barrier()
        nfr = max(12, min(args.frames * 2, 48))
        t0 = torch.cuda.Event(enable_timing=True)
        t1s = [torch.cuda.Event(enable_timing=True) for _ in lane_streams]
        t0.record(lane_streams[0])
        for s in lane_streams[1:]:
            s.wait_event(t0)  # no lane starts before the start mark
        for i in range(nfr):
            fn(i)
        for s, e in zip(lane_streams, t1s):
            e.record(s)
        sync_lanes(); barrier()
        te = torch.tensor([max(t0.elapsed_time(e) for e in t1s)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_val = wl.out_px * nfr * world / (float(te.item()) * 1e-3) / 1e6
        e2e_out = {"value": round(e2e_val, 1), "unit": "MP/s", "h2d_bytes_per_step": h2d * args.frames, "d2h_bytes_per_step": d2h * args.frames,
                   "note": "per frame: pinned host -> device (2 layers), kernel, device -> pinned host; 3 streams round robin so "
                           "copies overlap the kernels (PCIe bound)",
                   "host_affinity": ("%d CPUs of the GPU's NUMA node" % len(numa_cpus)) if numa_cpus else "unbound"}
        for c in lane_ctxs:
            c.close()
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    if rank == 0:
        peak, peak_src = hbm_peak()
        ms_step = total_ms_max / args.steps
        value = wl.out_px * wl.frames * world / (ms_step * 1e-3) / 1e6
        med = per[len(per) // 2]
        avg = total_ms / args.steps
        achieved = wl.bytes_per_frame * wl.frames / (avg * 1e-3) / 1e9
        traffic = None  # from the committed ncu capture of this very command (profiles/traffic.json)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(wl.name)
            if tj and tj["frames"] == wl.frames:
                traffic = tj["bytes"]
        except Exception:
            pass
        line = {
            "metric": "megapixels/sec", "value": round(value, 1), "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "what": wl.desc, "frames_per_step_per_gpu": wl.frames,
                       "out_px_per_frame": wl.out_px, "l2": "each step touches %.0f MB once (%s)" % (wl.bytes_per_frame * wl.frames / 1e6, "> 126 MB L2" if wl.bytes_per_frame * wl.frames > 126e6
                                                                         else "NOT larger than the 126 MB L2: raise --frames for a valid number"),
                       "parallelism": "frame-batch sharding over %d GPU(s), no collective" % world},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": wl.bytes_per_frame * wl.frames,
                         "kernel_ms_avg": round(avg, 4), "kernel_ms_median": round(med, 4)},
            "gpu_launches": int(launches),
            "clocks": sampler.result() if sampler else None,
        }
        if e2e_out:
            line["e2e"] = e2e_out
        if not args.no_cpu:
            cb = cpu_baseline(wl.name)
            if cb:
                line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
