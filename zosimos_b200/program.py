"""program / run / pool -- the compile -> lower -> launch -> step -> retire flow of the reference
(/root/reference/lib/zosimos/src/program.rs:1304, run.rs:471-1481, pool.rs:19-26) over the C-ABI.

    plan = Linker.from_included().compile(commands)        # Program
    executable = plan.lower_to(Capabilities.from_device(ctx))
    env = executable.from_pool(pool); env.bind(reg, key); env.recover_buffers()
    execution = executable.launch(env)
    while execution.is_running(): execution.step().block_on()
    retire = execution.retire_gracefully(pool); key = retire.output(out_reg).key(); retire.finish()

or, without an Executable (tests/direct.rs):  plan.launch(pool).bind(reg, key).launch(ctx)

What is re-used between launches (the reference's "re-use of the pipeline", Readme.md; tests/loop.rs):
  * the planned, fused kernel schedule (a `zos_program`) is cached by the Executable and taken again by the next
    launch on the same device -- the reference keeps its lowered instruction stream in `run::Executable` the same way;
  * device memory comes from the context's arena (zos_buf_alloc): `Retire.retire_buffers` / `finish` park the program's
    temporaries there, `Environment.recover_buffers` takes them back, `Pool.clear_cache` returns parked blocks to the
    driver.  After the first launch a relaunch performs no cudaMalloc;
  * pool images can live on the device (`Pool.upload`, ImageData::GpuBuffer of pool.rs:122-156): bound as inputs they are
    read in place, bound as outputs (`Environment.bind_output`) the program writes into them and nothing is downloaded.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import _ffi
from .buffer import Block, ByteLayout, Color, Descriptor
from .command import Register, RegisterKnob, descriptor_from_ffi, host_lib
from .device import Context, DeviceImage, PinnedArray


class LaunchError(Exception):  # program.rs:1997-2006
    pass


class StartError(Exception):  # run.rs:370-392
    pass


class StepError(Exception):  # run.rs:395-408
    pass


class RetireError(Exception):  # run.rs:429-448
    pass


class ImageUploadError(Exception):  # pool.rs:166-175
    pass


@dataclass(frozen=True)
class PoolKey:
    index: int


def host_nbytes(desc: Descriptor) -> int:
    """Bytes of one frame in the tight host layout: rows of width * texel_stride; planar 4:2:0 then its chroma planes."""
    w, h = desc.size()
    if desc.texel.block == Block.Pixel:
        return w * h * desc.layout.texel_stride
    cw, ch = (w + 1) // 2, (h + 1) // 2
    return w * h + 2 * cw * ch


class _Entry:
    """One pool item: descriptor + where the bytes are (pool.rs:122-156 ImageData::{Host, GpuBuffer, LateBound})."""
    __slots__ = ("desc", "host", "pinned", "device", "batch")

    def __init__(self, desc, host=None, pinned=None, device=None, batch=1):
        self.desc, self.host, self.pinned, self.device, self.batch = desc, host, pinned, device, batch

    def host_ptr(self) -> int:
        return self.pinned.ptr.value if self.pinned is not None else self.host.ctypes.data


class PoolImage:
    """pool.rs:51-66: a handle to one image of the pool."""

    def __init__(self, pool: "Pool", key: PoolKey):
        self._pool, self._key = pool, key

    def _e(self) -> _Entry:
        return self._pool._images[self._key.index]

    def key(self) -> PoolKey:
        return self._key

    def descriptor(self) -> Descriptor:
        return self._e().desc

    def layout(self) -> ByteLayout:
        return self.descriptor().layout

    def as_bytes(self) -> Optional[np.ndarray]:
        """pool.rs:669-676: None unless the data is on the host."""
        return self._e().host

    def is_device(self) -> bool:
        """ImageData::GpuBuffer (pool.rs:122-156)."""
        return self._e().device is not None

    def set_color(self, color: Color):
        e = self._e()
        e.desc = e.desc.with_color(color)

    def to_image(self) -> Optional[np.ndarray]:
        """(h, w, channels) view for 8-bit texels (the `image` crate conversion of pool.rs:640-648); None for device data."""
        e = self._e()
        if e.host is None:
            return None
        d = e.desc
        return e.host[: host_nbytes(d)].reshape(d.layout.height, d.layout.width, d.layout.texel_stride)


class Pool:
    """pool.rs:19-26: images plus devices.  Images are host byte arrays (optionally page-locked) or device buffers; the
    cache of the reference's pool (buffers, textures, shaders, pipelines parked between runs, pool.rs:93-99) is the
    device arena of each context plus the programs each Executable keeps."""

    def __init__(self, pin_host: bool = False):
        self._images: List[_Entry] = []
        self._devices: List[Context] = []
        self.pin_host = pin_host  # host images in page-locked memory: uploads / downloads are asynchronous and run at PCIe speed

    # -- devices (pool.rs:191-240)
    def request_device(self, index: int = 0) -> Context:
        ctx = Context(index)
        self._devices.append(ctx)
        return ctx

    def iter_devices(self):
        return iter(self._devices)

    # -- images (pool.rs:244-368)
    def _host_array(self, nbytes: int):
        if self.pin_host and self._devices:
            pin = PinnedArray(self._devices[0], nbytes)
            return pin.array, pin
        return np.empty(nbytes, np.uint8), None

    def insert(self, desc: Descriptor, data, batch: int = 1) -> PoolImage:
        a = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        if not desc.is_consistent():
            raise ValueError("inconsistent descriptor")
        if a.size != host_nbytes(desc) * batch:
            raise ValueError("data does not match the layout (%d bytes, expected %d)" % (a.size, host_nbytes(desc) * batch))
        host, pin = self._host_array(a.size)
        host[:] = a
        self._images.append(_Entry(desc, host, pin, None, batch))
        return PoolImage(self, PoolKey(len(self._images) - 1))

    def insert_srgb(self, rgba: np.ndarray) -> PoolImage:
        """pool.rs:253-265 for an RGBA8 image array (h, w, 4)."""
        h, w, c = rgba.shape
        assert c == 4 and rgba.dtype == np.uint8
        return self.insert(Descriptor.with_srgb_image("rgba8", w, h), rgba)

    def declare(self, desc: Descriptor, batch: int = 1) -> PoolImage:
        """pool.rs:281-290: an image without data (ImageData::LateBound)."""
        self._images.append(_Entry(desc, None, None, None, batch))
        return PoolImage(self, PoolKey(len(self._images) - 1))

    def allocate_like(self, key: PoolKey) -> PoolImage:
        e = self._images[key.index]
        if e.host is not None:
            return self.insert(e.desc, e.host, e.batch)
        host, pin = self._host_array(host_nbytes(e.desc) * e.batch)
        self._images.append(_Entry(e.desc, host, pin, None, e.batch))
        return PoolImage(self, PoolKey(len(self._images) - 1))

    def entry(self, key: PoolKey) -> Optional[PoolImage]:
        return PoolImage(self, key) if 0 <= key.index < len(self._images) else None

    def upload(self, key: PoolKey, ctx: Context):
        """pool.rs:292-430: move the image onto a device (ImageData::GpuBuffer: the aligned layout in a device buffer).
        A declared image without data becomes an uninitialised device image (a render target)."""
        if not (0 <= key.index < len(self._images)):
            raise ImageUploadError("BadImage")
        if ctx not in self._devices:
            raise ImageUploadError("BadGpu")
        e = self._images[key.index]
        if e.device is not None and e.device.ctx is ctx:
            return
        dev = ctx.image(e.desc, e.batch)
        if e.device is not None:  # on another device: through the host
            tmp = np.empty(host_nbytes(e.desc) * e.batch, np.uint8)
            e.device.download_into(tmp.ctypes.data)
            dev.upload_from(tmp.ctypes.data)
            e.device.free()
        elif e.host is not None:
            dev.upload_from(e.host_ptr())
        e.device, e.host = dev, None
        if e.pinned is not None:
            e.pinned.free()
            e.pinned = None

    def download(self, key: PoolKey):
        """The way back (the reference reads device images by running a copy program into a host image): device -> host."""
        e = self._images[key.index]
        if e.device is None:
            return
        host, pin = self._host_array(host_nbytes(e.desc) * e.batch)
        e.device.download_into(pin.ptr.value if pin is not None else host.ctypes.data)
        e.device.free()
        e.device, e.host, e.pinned = None, host, pin

    def clear_cache(self):
        """pool.rs:450-455: drop everything parked for re-use -- here the parked blocks of every device arena."""
        for ctx in self._devices:
            if ctx.handle:
                ctx.arena_trim()


@dataclass(frozen=True)
class Capabilities:
    """program.rs:367-370: here the device (ordinal, or a specific context of the pool), how registers are fused, and
    how many frames one launch processes (`batch` > 1: every register is a stack of `batch` frames)."""
    device: int = 0
    fuse_mode: int = _ffi.FUSE_EXACT
    batch: int = 1
    context: Optional[Context] = field(default=None, compare=False)

    @staticmethod
    def from_device(ctx: Context, fuse_mode: int = _ffi.FUSE_EXACT, batch: int = 1) -> "Capabilities":
        return Capabilities(ctx.device, fuse_mode, batch, ctx)


@dataclass(frozen=True)
class Knob:
    index: int


@dataclass
class StepLimits:
    """run.rs:223-225, 3001-3017."""
    instructions: int = 1

    @staticmethod
    def new() -> "StepLimits":
        return StepLimits(1)

    def with_steps(self, instructions: int) -> "StepLimits":
        return StepLimits(int(instructions))


@dataclass
class RecoveredBufferStats:  # run.rs:454-459
    mem: int = 0        # bytes taken over from parked blocks
    allocated: int = 0  # bytes that needed a fresh device allocation


@dataclass
class RetiredBufferStats:  # run.rs:461-469
    mem: int = 0
    buffer_keys: int = 0  # number of buffers parked (the reference lists their pool keys)


class Program:
    """program.rs:52-87: the linked High-level instruction stream."""

    def __init__(self, handle):
        self._h = handle

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

    # -- the direct interface (program.rs:1284-1302, 1030-1050, 1758; tests/direct.rs)
    @staticmethod
    def choose_adapter(adapters):
        """program.rs:1090-1117: the first usable adapter -- here a Context (or device ordinal) with sm_100 code."""
        for a in adapters:
            return a
        raise LaunchError("No matching adapter for program and adapter options")  # MismatchError

    def launch(self, pool: Pool) -> "Launcher":
        return Launcher(self, pool)


class Launcher:
    """program.rs:1030-1050, 1719-1800: bind images to a program and run it on an adapter, without an Executable."""

    def __init__(self, program: Program, pool: Pool):
        self.program, self.pool = program, pool
        self._binds: List = []
        self._outputs: List = []

    def bind(self, reg: Register, key: PoolKey) -> "Launcher":
        if self.pool.entry(key) is None:
            raise LaunchError("InternalCommandError: no such pool image")
        self._binds.append((reg, key))
        return self

    def bind_output(self, reg: Register, key: PoolKey) -> "Launcher":
        self._outputs.append((reg, key))
        return self

    def bind_remaining_outputs(self) -> "Launcher":  # program.rs:1735-1755: outputs get pool images when they retire
        return self

    def launch(self, adapter) -> "Execution":
        """`adapter`: a Context of the pool, or a device ordinal (a context is requested from the pool, like
        `adapter.request_device` in program.rs:1758-1762)."""
        ctx = adapter if isinstance(adapter, Context) else self.pool.request_device(int(adapter))
        if ctx not in self.pool._devices:
            self.pool._devices.append(ctx)
        exe = self.program.lower_to(Capabilities.from_device(ctx))
        env = exe.from_pool(self.pool)
        try:
            for reg, key in self._binds:
                env.bind(reg, key)
            for reg, key in self._outputs:
                env.bind_output(reg, key)
            return exe.launch(env)
        except StartError as e:
            raise LaunchError(str(e))


class Executable:
    """run.rs:38-60: the lowered program, re-launchable.  The planned kernel schedule (`zos_program`) of a finished
    execution is kept and handed to the next launch on the same context."""

    def __init__(self, program: Program, caps: Capabilities):
        self.program, self.caps = program, caps or Capabilities()
        self._ops = program.ops()
        self._inputs = {o.dst: o for o in self._ops if o.kind == _ffi.OP_INPUT}
        self._outputs = {o.reg: o for o in self._ops if o.kind == _ffi.OP_OUTPUT}
        self._idle: Dict[int, List] = {}  # id(ctx) -> planned programs waiting for the next launch
        self._ctxs: Dict[int, Context] = {}
        self.lowered = 0  # how many times a schedule was planned (1 per context unless executions overlap)

    def query_knob(self, knob: RegisterKnob) -> Optional[Knob]:
        k = int(host_lib().zosh_program_knob(self.program._h, int(knob.link_idx), int(knob.register.index)))
        return Knob(k) if k else None

    def from_pool(self, pool: Pool) -> "Environment":
        ctx = self.caps.context if self.caps.context in pool._devices else None
        if ctx is None:
            ctx = next((c for c in pool.iter_devices() if c.device == self.caps.device), None)
        if ctx is None:
            raise StartError("no device found in pool")
        return Environment(self, pool, ctx)

    def launch(self, env: "Environment") -> "Execution":
        return Execution(self, env)

    # -- the cache of planned programs
    def _acquire(self, ctx: Context):
        idle = self._idle.get(id(ctx))
        if idle:
            return idle.pop()
        h = C.c_void_p()
        st = host_lib().zosh_program_lower(self.program._h, ctx.handle, self.caps.fuse_mode, self.caps.batch, C.byref(h))
        if st != _ffi.OK:
            raise LaunchError((ctx._lib.zos_last_error(ctx.handle) or b"").decode())
        self.lowered += 1
        self._ctxs[id(ctx)] = ctx
        return h

    def _park(self, ctx: Context, prog):
        self._idle.setdefault(id(ctx), []).append(prog)

    def close(self):
        for cid, progs in self._idle.items():
            ctx = self._ctxs.get(cid)
            if ctx is not None and ctx.handle:
                for p in progs:
                    ctx._lib.zos_program_destroy(p)
        self._idle.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Environment:
    """run.rs:88-120: bindings of inputs / outputs / knobs for one launch."""

    def __init__(self, exe: Executable, pool: Pool, ctx: Context):
        self.exe, self.pool, self.ctx = exe, pool, ctx
        self.inputs: Dict[int, PoolKey] = {}
        self.outputs: Dict[int, PoolKey] = {}
        self.knobs: Dict[int, bytes] = {}
        self._prog = None
        self._order: List = []  # device allocations of this launch, in order ("T" = the program's temporaries)

    def _bindable(self, key: PoolKey, want: Descriptor, what: str) -> _Entry:
        img = self.pool.entry(key)
        if img is None:
            raise StartError("InternalCommandError: no such pool image")
        e = img._e()
        have = e.desc
        if (want.size(), want.texel) != (have.size(), have.texel):  # layout + texel; colour semantics ignored (run.rs:1184-1191)
            raise StartError("MismatchedDescriptor for %s" % what)  # run.rs:376-380
        if e.batch != self.exe.caps.batch:
            raise StartError("MismatchedDescriptor for %s: %d frames bound to a program of batch %d" % (what, e.batch, self.exe.caps.batch))
        return e

    def bind(self, reg: Register, key: PoolKey):
        idx = self.exe.program.register_index(reg.index)
        op = self.exe._inputs.get(idx)
        if op is None:
            raise StartError("register %d is not an input (StartError::MissingKey)" % reg.index)
        e = self._bindable(key, descriptor_from_ffi(op.desc), "input register %d" % reg.index)
        if e.host is None and e.device is None:
            raise StartError("pool image without data (ImageData::LateBound)")  # run.rs:1193-1198
        self.inputs[idx] = key

    def bind_output(self, reg: Register, key: PoolKey):
        """run.rs:1207-1242: the output lands in this pool image -- written in place when it lives on this device."""
        idx = self.exe.program.register_index(reg.index)
        op = self.exe._outputs.get(idx)
        if op is None:
            raise StartError("register %d is not an output (StartError::MissingKey)" % reg.index)
        self._bindable(key, descriptor_from_ffi(op.desc), "output register %d" % reg.index)
        self.outputs[idx] = key

    def bind_render(self, reg: Register, key: PoolKey):
        """run.rs:1244-1281: like bind_output, for a target that already lives on the device."""
        img = self.pool.entry(key)
        if img is None or not img.is_device():
            raise StartError("InternalCommandError: a render target must be a device image")
        self.bind_output(reg, key)

    def knob(self, knob: Knob, data: bytes):
        self.knobs[knob.index] = bytes(data)

    def knob_by_register(self, knob: RegisterKnob, data: bytes):  # run.rs:1283-1290
        k = self.exe.query_knob(knob)
        if k is None:
            raise KeyError("Knob does not exist in this program")
        self.knob(k, data)

    def _program(self):
        if self._prog is None:
            self._prog = self.exe._acquire(self.ctx)
        return self._prog

    def recover_buffers(self) -> RecoveredBufferStats:
        """run.rs:1312-1347: take matching temporaries of earlier runs out of the cache (the context's arena)."""
        reused, fresh = C.c_uint64(0), C.c_uint64(0)
        self.ctx.check(self.ctx._lib.zos_program_recover_buffers(self._program(), C.byref(reused), C.byref(fresh)))
        if "T" not in self._order:
            self._order.append("T")
        return RecoveredBufferStats(int(reused.value), int(fresh.value))


class SyncPoint:
    def __init__(self, ctx: Context):
        self._ctx = ctx

    def block_on(self):
        self._ctx.sync()

    async def finish(self, queue_poll=None):
        """run.rs:3041-3169 (tests/async.rs): step towards the synchronisation point without blocking the thread.  `queue_poll`, if given, is
        called with the device context and may return a guard object that is dropped when the wait is over (the reference's callers
        schedule their device polling there; CUDA needs none -- the stream is queried between yields to the event loop)."""
        import asyncio
        guard = queue_poll(self._ctx) if queue_poll is not None else None
        try:
            while not self._ctx.poll():
                await asyncio.sleep(0)
        finally:
            del guard


class Execution:
    """run.rs:211-260."""

    def __init__(self, exe: Executable, env: Environment):
        self.exe, self.env, self.ctx = exe, env, env.ctx
        lib = self.ctx._lib
        self._staged: Dict[int, DeviceImage] = {}
        self._order = env._order
        self._prog = env._program()
        env._prog = None
        self._done = False
        try:
            for idx, op in exe._inputs.items():
                key = env.inputs.get(idx)
                if key is None:  # unused inputs may stay unbound; a needed one fails in zos_program_launch
                    self._check(lib.zos_program_unbind(self._prog, idx), StartError)  # (a cached program remembers its last bindings)
                    continue
                e = env.pool._images[key.index]
                if e.device is not None and e.device.ctx is self.ctx:
                    dev = e.device  # ImageData::GpuBuffer: read in place
                else:
                    if e.host is None:  # on another device
                        env.pool.download(key)
                    dev = self.ctx.image(e.desc, e.batch)
                    self._staged[idx] = dev
                    self._order.append(dev)
                    dev.upload_from(e.host_ptr(), sync=e.pinned is None)  # pinned: asynchronous on the context's stream
                im = dev.ffi()
                self._check(lib.zos_program_bind(self._prog, idx, C.byref(im)), StartError)
            for idx, op in exe._outputs.items():
                key = env.outputs.get(idx)
                e = env.pool._images[key.index] if key is not None else None
                if op.src[0] in exe._inputs:
                    continue  # an input handed through: it is the input's binding
                if e is not None and e.device is not None and e.device.ctx is self.ctx:
                    im = e.device.ffi()
                    self._check(lib.zos_program_bind(self._prog, op.src[0], C.byref(im)), StartError)
                else:
                    self._check(lib.zos_program_unbind(self._prog, op.src[0]), StartError)
            self._check(lib.zos_program_reset_knobs(self._prog), StartError)  # (a cached program remembers the last environment's knobs)
            for k, data in env.knobs.items():
                buf = C.create_string_buffer(data, len(data))
                self._check(lib.zos_program_set_knob(self._prog, k, buf, len(data)), StartError)
            if "T" not in self._order:
                self._order.append("T")  # zos_program_launch takes the temporaries back if they were parked
            self._check(lib.zos_program_launch(self._prog), StartError)
        except Exception:
            self._discard()
            raise
        self._running = lib.zos_program_kernel_count(self._prog) > 0

    def _check(self, st, exc):
        if st != _ffi.OK:
            raise exc((self.ctx._lib.zos_last_error(self.ctx.handle) or b"").decode())

    def _free_in_reverse(self, park: bool):
        """Frees what this launch allocated in the reverse order, so that the arena's most-recently-parked-first lists
        hand the same blocks to the same users next time (a captured CUDA graph then stays valid)."""
        lib = self.ctx._lib
        for item in reversed(self._order):
            if item == "T":
                if park and self._prog:
                    lib.zos_program_release_buffers(self._prog, None, None)
            else:
                item.free()
        self._order.clear()
        self._staged.clear()

    def _discard(self):
        if self._done:
            return
        self._done = True
        if self.ctx.handle:
            self._free_in_reverse(park=False)
            if self._prog:
                self.ctx._lib.zos_program_destroy(self._prog)
        self._prog = None

    def __del__(self):
        try:
            self._discard()
        except Exception:
            pass

    def kernel_count(self) -> int:
        return int(self.ctx._lib.zos_program_kernel_count(self._prog))

    def is_running(self) -> bool:
        return self._running

    def step(self) -> SyncPoint:
        return self.step_to(StepLimits(1))

    def step_to(self, limits: StepLimits) -> SyncPoint:
        """run.rs:1394-1460: up to `limits.instructions` launches of the schedule."""
        if not self._running:
            raise StepError("ProgramEnd")
        r = C.c_int32(0)
        self._check(self.ctx._lib.zos_program_step(self._prog, max(int(limits.instructions), 1), C.byref(r)), StepError)
        self._running = bool(r.value)
        return SyncPoint(self.ctx)

    def rerun(self, knobs: Optional[Dict["Knob", bytes]] = None, graph: bool = True) -> SyncPoint:
        """Run the same plan again with the same bindings, optionally with other knob values, as one CUDA-graph
        submission (`zos_program_run`; tests/knobs.rs)."""
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

    def resources_used(self) -> dict:
        """run.rs:1477-1480 `ResourcesUsed`: what this execution holds, and what the device arena did so far."""
        st = _ffi.ZosProgramStats()
        self._check(self.ctx._lib.zos_program_resources(self._prog, C.byref(st)), StepError)
        out = {k: int(getattr(st, k)) for k, _ in _ffi.ZosProgramStats._fields_ if k != "reserved"}
        out["arena"] = self.ctx.arena_stats()
        return out

    def retire(self):
        """run.rs:1448-1458: stop, discarding every resource."""
        self.retire_gracefully(Pool()).finish_by_discarding()

    def retire_gracefully(self, pool: Pool) -> "Retire":
        if self._running:
            raise RetireError("execution is still running")
        return Retire(self, pool)


class Retire:
    """run.rs:2786-2997: moves results back into the pool."""

    def __init__(self, execution: Execution, pool: Pool):
        self.ex, self.pool = execution, pool

    def input(self, reg: Register) -> PoolImage:
        """run.rs:2791-2806: the input image goes back to the pool (it never left it here)."""
        idx = self.ex.exe.program.register_index(reg.index)
        key = self.ex.env.inputs.get(idx)
        if idx not in self.ex.exe._inputs or key is None:
            raise RetireError("NoSuchInput")
        return PoolImage(self.pool, key)

    def output_key(self, reg: Register) -> Optional[PoolKey]:  # run.rs:2860-2873
        idx = self.ex.exe.program.register_index(reg.index)
        if idx not in self.ex.exe._outputs:
            raise RetireError("NoSuchOutput")
        return self.ex.env.outputs.get(idx)

    def output(self, reg: Register) -> PoolImage:
        ex = self.ex
        idx = ex.exe.program.register_index(reg.index)
        op = ex.exe._outputs.get(idx)
        if op is None:
            raise RetireError("NoSuchOutput: register %d" % reg.index)
        key = ex.env.outputs.get(idx)
        entry = self.pool._images[key.index] if key is not None and self.pool.entry(key) is not None else None
        if entry is not None and entry.device is not None and entry.device.ctx is ex.ctx and op.src[0] not in ex.exe._inputs:
            return PoolImage(self.pool, key)  # the program wrote into the pool's device image
        im = _ffi.ZosImage()
        if ex.ctx._lib.zos_program_register_image(ex._prog, op.src[0], C.byref(im)) != _ffi.OK:
            raise RetireError("output register has no storage")
        desc = descriptor_from_ffi(im.desc)
        batch, fb = ex.exe.caps.batch, host_nbytes(desc)  # planar frames: Y plus both chroma planes
        if entry is None:
            host, pin = self.pool._host_array(fb * batch)
            self.pool._images.append(_Entry(desc, host, pin, None, batch))
            key = PoolKey(len(self.pool._images) - 1)
            entry = self.pool._images[key.index]
        elif entry.host is None or entry.host.size != fb * batch:
            if entry.device is not None:
                entry.device.free()
                entry.device = None
            entry.host, entry.pinned = self.pool._host_array(fb * batch)
        ptr = entry.host_ptr()
        for f in range(batch):
            ex.ctx.check(ex.ctx._lib.zos_image_download(ex.ctx.handle, C.byref(im), f, C.c_void_p(ptr + f * fb)))
        ex.ctx.sync()
        return PoolImage(self.pool, key)

    render = output  # run.rs:2826-2840

    def retire_buffers(self) -> RetiredBufferStats:
        """run.rs:2876-2942: the temporaries of this run are kept for the next one.  They reach the device arena when the
        retirement finishes (`finish` parks everything in the reverse order of its allocation, see
        Execution._free_in_reverse); this call reports what will be parked."""
        st = _ffi.ZosProgramStats()
        ex = self.ex
        if not ex._prog:
            return RetiredBufferStats()
        ex.ctx.check(ex.ctx._lib.zos_program_resources(ex._prog, C.byref(st)))
        return RetiredBufferStats(int(st.temp_bytes), int(st.temp_buffers))

    def prune(self):  # run.rs:2944-2952 (not implemented there either)
        pass

    def finish(self):
        """run.rs:2954-2986: everything that can be used again stays -- temporaries in the arena, the planned program
        with its executable."""
        ex = self.ex
        if ex._done:
            return
        ex._done = True
        ex._free_in_reverse(park=True)
        if ex._prog:
            ex.exe._park(ex.ctx, ex._prog)
            ex._prog = None

    def finish_by_discarding(self):  # run.rs:2988-2996
        self.ex._discard()
