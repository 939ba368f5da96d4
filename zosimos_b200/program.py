"""program / run / pool -- the compile -> lower -> launch -> step -> retire flow of the reference
(/root/reference/lib/zosimos/src/program.rs:1304, run.rs:471-1481, pool.rs:19-26) over the C-ABI.

    plan = Linker.from_included().compile(commands)        # Program
    executable = plan.lower_to(Capabilities.from_pool(pool))
    env = executable.from_pool(pool); env.bind(reg, key)
    execution = executable.launch(env)
    while execution.is_running(): execution.step().block_on()
    retire = execution.retire_gracefully(pool); key = retire.output(out_reg).key(); retire.finish()

Host images live in the `Pool` as tight-row byte arrays; launching uploads the bound inputs into
256-byte-pitched device buffers (buffer.rs:121-134), retiring downloads the outputs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import _ffi
from .buffer import Block, ByteLayout, Color, Descriptor, SampleParts, Texel
from .command import CommandError, Register, RegisterKnob, descriptor_from_ffi, host_lib
from .device import Context, DeviceImage


class LaunchError(Exception):  # program.rs:1997-2006
    pass


class StartError(Exception):  # run.rs:370-392
    pass


class StepError(Exception):  # run.rs:395-408
    pass


class RetireError(Exception):  # run.rs:429-448
    pass


@dataclass(frozen=True)
class PoolKey:
    index: int


class PoolImage:
    """pool.rs:51-66: a handle to one image of the pool."""

    def __init__(self, pool: "Pool", key: PoolKey):
        self._pool, self._key = pool, key

    def key(self) -> PoolKey:
        return self._key

    def descriptor(self) -> Descriptor:
        return self._pool._images[self._key.index][0]

    def layout(self) -> ByteLayout:
        return self.descriptor().layout

    def as_bytes(self) -> Optional[np.ndarray]:
        return self._pool._images[self._key.index][1]

    def set_color(self, color: Color):
        d, data = self._pool._images[self._key.index]
        self._pool._images[self._key.index] = (d.with_color(color), data)

    def to_image(self) -> np.ndarray:
        """(h, w, channels) view for 8-bit texels (the `image` crate conversion of pool.rs)."""
        d = self.descriptor()
        return self.as_bytes().reshape(d.layout.height, d.layout.width, d.layout.texel_stride)


class Pool:
    """pool.rs:19-26: images plus devices.  Device-side caches of the reference (textures, pipelines,
    shaders) have no equivalent: programs own their device buffers."""

    def __init__(self):
        self._images: List = []
        self._devices: List[Context] = []

    # -- devices (pool.rs:205-240)
    def request_device(self, index: int = 0) -> Context:
        ctx = Context(index)
        self._devices.append(ctx)
        return ctx

    def iter_devices(self):
        return iter(self._devices)

    # -- images (pool.rs:244-368)
    def insert(self, desc: Descriptor, data) -> PoolImage:
        a = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        if not desc.is_consistent():
            raise ValueError("inconsistent descriptor")
        if desc.texel.block == Block.Pixel and a.size != desc.layout.height * desc.layout.width * desc.layout.texel_stride:
            raise ValueError("data does not match the layout")
        self._images.append((desc, a.copy()))
        return PoolImage(self, PoolKey(len(self._images) - 1))

    def insert_srgb(self, rgba: np.ndarray) -> PoolImage:
        """pool.rs:262-270 for an RGBA8 image array (h, w, 4)."""
        h, w, c = rgba.shape
        assert c == 4 and rgba.dtype == np.uint8
        return self.insert(Descriptor.with_srgb_image("rgba8", w, h), rgba)

    def declare(self, desc: Descriptor) -> PoolImage:
        self._images.append((desc, None))
        return PoolImage(self, PoolKey(len(self._images) - 1))

    def allocate_like(self, key: PoolKey) -> PoolImage:
        d, data = self._images[key.index]
        return self.insert(d, data)

    def entry(self, key: PoolKey) -> Optional[PoolImage]:
        return PoolImage(self, key) if 0 <= key.index < len(self._images) else None

    def clear_cache(self):
        pass


@dataclass(frozen=True)
class Capabilities:
    """program.rs:367-370: here simply the device ordinal (and how registers are fused)."""
    device: int = 0
    fuse_mode: int = _ffi.FUSE_EXACT

    @staticmethod
    def from_device(ctx: Context, fuse_mode: int = _ffi.FUSE_EXACT) -> "Capabilities":
        return Capabilities(ctx.device, fuse_mode)


@dataclass(frozen=True)
class Knob:
    index: int


class Program:
    """program.rs:52-87: the linked High-level instruction stream."""

    def __init__(self, handle, knobs: Dict[int, int]):
        self._h = handle
        self._knobs = knobs

    def __del__(self):
        try:
            if self._h:
                host_lib().zosh_program_free(self._h)
                self._h = None
        except Exception:
            pass

    def ops(self) -> List[_ffi.ZosOp]:
        n = host_lib().zosh_program_num_ops(self._h)
        p = host_lib().zosh_program_ops(self._h)
        # copies: p[i] would be a view into the program's memory, dangling once a temporary Program is collected
        # (the data / source pointers inside still belong to the program -- keep it alive while they are used)
        return [_ffi.ZosOp.from_buffer_copy(p[i]) for i in range(n)]

    def register_index(self, index: int) -> int:
        """Register of the command buffer that was linked -> register of this program: the identity except for a generic
        entry point, whose program is its monomorphic copy (zosh_program_register)."""
        return int(host_lib().zosh_program_register(self._h, int(index)))

    def lower_to(self, capabilities: Capabilities) -> "Executable":
        return Executable(self, capabilities)


class Executable:
    """run.rs:38: re-launchable; every launch builds a zos_program on the device of `capabilities`."""

    def __init__(self, program: Program, caps: Capabilities):
        self.program, self.caps = program, caps
        self._ops = program.ops()

    def query_knob(self, knob: RegisterKnob) -> Optional[Knob]:
        k = self.program._knobs.get(knob.register.index)
        return Knob(k) if k else None

    def from_pool(self, pool: Pool) -> "Environment":
        ctx = next((c for c in pool.iter_devices() if c.device == self.caps.device), None)
        if ctx is None:
            raise StartError("no device found in pool")
        return Environment(self, pool, ctx)

    def launch(self, env: "Environment") -> "Execution":
        return Execution(self, env)


class Environment:
    """run.rs:88: bindings of inputs / outputs / knobs for one launch."""

    def __init__(self, exe: Executable, pool: Pool, ctx: Context):
        self.exe, self.pool, self.ctx = exe, pool, ctx
        self.inputs: Dict[int, PoolKey] = {}
        self.knobs: Dict[int, bytes] = {}

    def bind(self, reg: Register, key: PoolKey):
        idx = self.exe.program.register_index(reg.index)
        op = next((o for o in self.exe._ops if o.kind == _ffi.OP_INPUT and o.dst == idx), None)
        if op is None:
            raise StartError("register %d is not an input (StartError::MissingKey)" % reg.index)
        img = self.pool.entry(key)
        if img is None or img.as_bytes() is None:
            raise StartError("pool key without host data")
        want = descriptor_from_ffi(op.desc)
        have = img.descriptor()
        if (want.size(), want.texel, want.color) != (have.size(), have.texel, have.color):
            raise StartError("MismatchedDescriptor for register %d" % reg.index)  # run.rs:376-380
        self.inputs[idx] = key

    def knob(self, knob: Knob, data: bytes):
        self.knobs[knob.index] = bytes(data)

    def recover_buffers(self):
        return None


class SyncPoint:
    def __init__(self, ctx: Context):
        self._ctx = ctx

    def block_on(self):
        self._ctx.sync()


class Execution:
    """run.rs:211-260."""

    def __init__(self, exe: Executable, env: Environment):
        self.exe, self.env, self.ctx = exe, env, env.ctx
        lib = self.ctx._lib
        self._device_inputs: Dict[int, DeviceImage] = {}
        h = C.c_void_p()
        st = host_lib().zosh_program_lower(exe.program._h, self.ctx.handle, exe.caps.fuse_mode, 1, C.byref(h))
        if st != _ffi.OK:
            raise LaunchError((lib.zos_last_error(self.ctx.handle) or b"").decode())
        self._prog = h
        try:
            for op in exe._ops:
                if op.kind != _ffi.OP_INPUT:
                    continue
                key = env.inputs.get(op.dst)
                if key is None:
                    continue  # unused inputs may stay unbound; a needed one fails in zos_program_launch
                d, data = env.pool._images[key.index]
                dev = self.ctx.image(d)
                if d.texel.block == Block.Pixel:
                    dev.upload(data)
                else:
                    w, hh = d.size()
                    cw, ch = (w + 1) // 2, (hh + 1) // 2
                    y = data[: w * hh]
                    if d.texel.block == Block.Yuv420Nv12:
                        dev.upload((y, data[w * hh: w * hh + 2 * cw * ch], None))
                    else:
                        dev.upload((y, data[w * hh: w * hh + cw * ch], data[w * hh + cw * ch: w * hh + 2 * cw * ch]))
                self._device_inputs[op.dst] = dev
                im = dev.ffi()
                self._check(lib.zos_program_bind(self._prog, op.dst, C.byref(im)), StartError)
            for k, data in env.knobs.items():
                buf = C.create_string_buffer(data, len(data))
                self._check(lib.zos_program_set_knob(self._prog, k, buf, len(data)), StartError)
            self._check(lib.zos_program_launch(self._prog), StartError)
        except Exception:
            lib.zos_program_destroy(self._prog)
            self._prog = None
            raise
        self._running = lib.zos_program_kernel_count(self._prog) > 0

    def _check(self, st, exc):
        if st != _ffi.OK:
            raise exc((self.ctx._lib.zos_last_error(self.ctx.handle) or b"").decode())

    def kernel_count(self) -> int:
        return int(self.ctx._lib.zos_program_kernel_count(self._prog))

    def is_running(self) -> bool:
        return self._running

    def step(self) -> SyncPoint:
        if not self._running:
            raise StepError("ProgramEnd")
        r = C.c_int32(0)
        self._check(self.ctx._lib.zos_program_step(self._prog, 1, C.byref(r)), StepError)
        self._running = bool(r.value)
        return SyncPoint(self.ctx)

    def rerun(self, knobs: Optional[Dict["Knob", bytes]] = None, graph: bool = True) -> SyncPoint:
        """Executable reuse (run.rs:1283-1347, tests/loop.rs, tests/knobs.rs): run the same plan again,
        optionally with other knob values, as one CUDA-graph submission (`zos_program_run`)."""
        if self._running:
            raise StepError("execution is still being stepped")
        lib = self.ctx._lib
        for k, data in (knobs or {}).items():
            buf = C.create_string_buffer(bytes(data), len(data))
            self._check(lib.zos_program_set_knob(self._prog, k.index, buf, len(data)), StartError)
        self._check(lib.zos_program_run(self._prog, 1 if graph else 0), StepError)
        return SyncPoint(self.ctx)

    def graph_launches(self) -> int:
        return int(self.ctx._lib.zos_program_graph_launches(self._prog))

    def retire_gracefully(self, pool: Pool) -> "Retire":
        if self._running:
            raise RetireError("execution is still running")
        return Retire(self, pool)


class Retire:
    """run.rs:2786-2997: moves results back into the pool."""

    def __init__(self, execution: Execution, pool: Pool):
        self.ex, self.pool = execution, pool

    def output(self, reg: Register) -> PoolImage:
        ex = self.ex
        idx = ex.exe.program.register_index(reg.index)
        op = next((o for o in ex.exe._ops if o.kind == _ffi.OP_OUTPUT and o.reg == idx), None)
        if op is None:
            raise RetireError("register %d is not an output" % reg.index)
        im = _ffi.ZosImage()
        if ex.ctx._lib.zos_program_register_image(ex._prog, op.src[0], C.byref(im)) != _ffi.OK:
            raise RetireError("output register has no storage")
        desc = descriptor_from_ffi(im.desc)
        w, h = desc.size()
        out = np.empty(h * w * desc.layout.texel_stride, np.uint8)
        ex.ctx.check(ex.ctx._lib.zos_image_download(ex.ctx.handle, C.byref(im), 0, out.ctypes.data_as(C.c_void_p)))
        ex.ctx.sync()
        return self.pool.insert(desc, out)

    def retire_buffers(self):
        return None

    def finish(self):
        lib = self.ex.ctx._lib
        if self.ex._prog:
            lib.zos_program_destroy(self.ex._prog)
            self.ex._prog = None
        for dev in self.ex._device_inputs.values():
            dev.free()
        self.ex._device_inputs.clear()
