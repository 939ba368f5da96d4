"""GPU tests of what is re-used between launches (the reference's `tests/loop.rs`, `tests/direct.rs`,
run.rs:1207-1347, 1394, 2876-2942, pool.rs:122-156, 292): the planned program cached by the Executable, the device arena
behind recover_buffers / retire_buffers (no cudaMalloc after the first launch), pool images that live on the device,
bound outputs, `Program::launch(pool)` -> `Launcher`, `step_to(StepLimits)`, batched programs."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.gpu_common import oracle_image

pytestmark = pytest.mark.gpu

import zosimos_b200 as Z  # noqa: E402
from zosimos_b200.buffer import Color, Descriptor, SampleParts, Texel, Transfer  # noqa: E402
from zosimos_b200.command import Blend, ChromaticAdaptationMethod, CommandBuffer, Linker, Rectangle  # noqa: E402
from zosimos_b200.program import (Capabilities, LaunchError, Pool, RetireError, StartError, StepError, StepLimits)  # noqa: E402


def hashes():
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_hashes.json")) as f:
        return json.load(f)


@pytest.fixture()
def pool():
    p = Pool()
    p.request_device(0)
    yield p
    for c in p.iter_devices():
        c.close()


def inscribe_program(bg_desc, fg_desc):
    c = CommandBuffer()
    background, foreground = c.input(bg_desc), c.input(fg_desc)
    result = c.inscribe(background, Rectangle(0, 0, fg_desc.layout.width, fg_desc.layout.height), foreground)
    output, _ = c.output(result)
    return c, background, foreground, output


def run_like_util_rs(executable, pool, binds, out_reg):
    """tests/util.rs:85-118, call for call."""
    env = executable.from_pool(pool)
    for reg, key in binds:
        env.bind(reg, key)
    recovered = env.recover_buffers()
    execution = executable.launch(env)
    pool.clear_cache()
    while execution.is_running():
        execution.step().block_on()
    retire = execution.retire_gracefully(pool)
    key = retire.output(out_reg).key()
    retired = retire.retire_buffers()
    retire.finish()
    return key, recovered, retired


def test_loop_rs_relaunch_allocates_nothing_after_the_first_launch(pool, fixtures):
    """tests/loop.rs:80-95: one Executable, relaunched with upload + read-back.  After launch 1 the device arena serves every
    allocation (inputs staged from the host, the output register) from parked blocks and the schedule is planned once."""
    bg, fg = pool.insert_srgb(fixtures["background"]), pool.insert_srgb(fixtures["foreground"])
    c, background, foreground, output = inscribe_program(bg.descriptor(), fg.descriptor())
    ctx = next(pool.iter_devices())
    executable = Linker.from_included().compile(c).lower_to(Capabilities.from_device(ctx))
    result = bg.key()
    stats = []
    for it in range(12):
        result, recovered, retired = run_like_util_rs(executable, pool, [(background, bg.key()), (foreground, fg.key())], output)
        stats.append((ctx.arena_stats(), recovered, retired))
    first, last = stats[0][0], stats[-1][0]
    assert executable.lowered == 1
    # util.rs clears the pool cache right after every launch, so what was parked by the previous run and not recovered is
    # handed back; the program's own register IS recovered (run.rs:1312-1347): no allocation for it after run 1
    assert stats[0][1].mem == 0 and stats[0][1].allocated == 0  # nothing was planned with parked storage before launch 1
    assert all(s[1].mem == 512 * 512 * 4 and s[1].allocated == 0 for s in stats[1:])
    assert all(s[2].mem == 512 * 512 * 4 and s[2].buffer_keys == 1 for s in stats)
    img = pool.entry(result)
    got = img.as_bytes().reshape(512, 512, 4)
    assert O.blockhash256(got) in hashes()["composed"]
    exp = O.inscribe(oracle_image(bg.descriptor(), fixtures["background"]), (0, 0, 157, 151), oracle_image(fg.descriptor(), fixtures["foreground"]))
    assert np.array_equal(img.as_bytes(), exp.data.reshape(-1))

    # the same loop WITHOUT clearing the cache (what a throughput loop does): zero cudaMalloc after launch 1
    base = ctx.arena_stats()["device_allocs"]
    for it in range(10):
        env = executable.from_pool(pool)
        env.bind(background, bg.key()); env.bind(foreground, fg.key())
        execution = executable.launch(env)
        while execution.is_running():
            execution.step()
        retire = execution.retire_gracefully(pool)
        key = retire.output(output).key()
        retire.finish()
        if it == 0:
            after_first = ctx.arena_stats()["device_allocs"]
    end = ctx.arena_stats()
    assert end["device_allocs"] == after_first, (base, after_first, end)
    assert end["reuses"] >= last["reuses"] + 9 * 3
    assert end["bytes_in_use"] == 0 and end["bytes_parked"] == end["bytes_reserved"]
    assert np.array_equal(pool.entry(key).as_bytes(), exp.data.reshape(-1))
    pool.clear_cache()
    assert ctx.arena_stats()["bytes_reserved"] == 0


def test_device_resident_pool_images_and_bound_outputs(pool, fixtures):
    """ImageData::GpuBuffer (pool.rs:122-156, 292): inputs that live on the device are read in place, an output bound to a
    device image (run.rs:1207-1242) is written in place; nothing crosses PCIe during the run."""
    bg, fg = pool.insert_srgb(fixtures["background"]), pool.insert_srgb(fixtures["foreground"])
    ctx = next(pool.iter_devices())
    c, background, foreground, output = inscribe_program(bg.descriptor(), fg.descriptor())
    executable = Linker.from_included().compile(c).lower_to(Capabilities.from_device(ctx))
    host_key, _, _ = run_like_util_rs(executable, pool, [(background, bg.key()), (foreground, fg.key())], output)
    expect = np.array(pool.entry(host_key).as_bytes(), copy=True)

    pool.upload(bg.key(), ctx); pool.upload(fg.key(), ctx)
    assert bg.is_device() and bg.as_bytes() is None and bg.to_image() is None
    target = pool.declare(bg.descriptor())
    pool.upload(target.key(), ctx)  # an uninitialised device image: a render target
    allocs = ctx.arena_stats()["device_allocs"]
    for _ in range(3):
        env = executable.from_pool(pool)
        env.bind(background, bg.key()); env.bind(foreground, fg.key())
        env.bind_render(output, target.key())
        execution = executable.launch(env)
        assert execution.resources_used()["temp_bytes"] == 0  # every register of this program is provided by the pool
        while execution.is_running():
            execution.step().block_on()
        retire = execution.retire_gracefully(pool)
        assert retire.output_key(output) == target.key()
        out = retire.output(output)
        assert out.key() == target.key() and out.is_device()
        assert retire.input(background).key() == bg.key()
        retire.finish()
    assert ctx.arena_stats()["device_allocs"] == allocs  # no staging, no temporaries
    pool.download(target.key())
    assert np.array_equal(target.as_bytes(), expect)
    # a host image bound as the output receives the bytes in place (no new pool entry)
    n = len(pool._images)
    host_target = pool.insert(bg.descriptor(), np.zeros(512 * 512 * 4, np.uint8))
    env = executable.from_pool(pool)
    env.bind(background, bg.key()); env.bind(foreground, fg.key()); env.bind_output(output, host_target.key())
    with pytest.raises(StartError):
        env.bind_render(output, host_target.key())  # a render target must live on the device (run.rs:1268-1273)
    with pytest.raises(StartError):
        env.bind_output(background, host_target.key())  # not an output
    execution = executable.launch(env)
    execution.step_to(StepLimits.new().with_steps(10)).block_on()
    retire = execution.retire_gracefully(pool)
    assert retire.output(output).key() == host_target.key() and len(pool._images) == n + 1
    with pytest.raises(RetireError):
        retire.output(background)
    retire.finish()
    assert np.array_equal(host_target.as_bytes(), expect)


def test_direct_launcher(pool, fixtures):
    """tests/direct.rs: Program::launch(pool).bind(..).bind(..).launch(adapter), stepped, retired."""
    bg, fg = pool.insert_srgb(fixtures["background"]), pool.insert_srgb(fixtures["foreground"])
    c, background, foreground, output = inscribe_program(bg.descriptor(), fg.descriptor())
    plan = Linker.from_included().compile(c)
    adapter = plan.choose_adapter(pool.iter_devices())
    execution = plan.launch(pool).bind(background, bg.key()).bind(foreground, fg.key()).launch(adapter)
    while execution.is_running():
        execution.step()
    retire = execution.retire_gracefully(pool)
    image = retire.output(output)
    assert O.blockhash256(image.as_bytes().reshape(512, 512, 4)) in hashes()["composed"]
    retire.finish()
    with pytest.raises(LaunchError):  # an input was never supplied (program.rs:1766-1772)
        plan.launch(pool).bind(background, bg.key()).launch(adapter)
    with pytest.raises(LaunchError):
        plan.launch(pool).bind(background, fg.key()).bind(foreground, fg.key()).launch(adapter)  # wrong layout


def test_step_to_limits_and_empty_schedules(pool, fixtures):
    """run.rs:1389-1460: step() is step_to(1 instruction); a larger limit launches several kernels per sync point.  An output
    taken straight from an input has nothing to step through and can be retired (and run again) at once."""
    bg = pool.insert_srgb(fixtures["background"])
    rgba8 = Texel.new_u8(SampleParts.RgbA)
    c = CommandBuffer()
    i = c.input(bg.descriptor())
    a = c.color_convert(i, Color.BT709_RGB, rgba8)
    o1, _ = c.output(a)                                   # `a` is read twice: it is materialised, two kernels
    o2, _ = c.output(c.chromatic_adaptation(a, ChromaticAdaptationMethod.VonKries, Z.Whitepoint.D50))
    executable = Linker.from_included().compile(c).lower_to(Capabilities.from_device(next(pool.iter_devices())))

    def launch():
        env = executable.from_pool(pool); env.bind(i, bg.key())
        return executable.launch(env)
    one = launch()
    assert one.kernel_count() == 2
    steps = 0
    while one.is_running():
        one.step().block_on(); steps += 1
    r1 = one.retire_gracefully(pool); by_one = [np.array(r1.output(o).as_bytes(), copy=True) for o in (o1, o2)]; r1.finish()
    many = launch()
    many.step_to(StepLimits(64)).block_on()
    assert steps == 2 and not many.is_running()
    with pytest.raises(StepError):
        many.step_to(StepLimits(1))
    r2 = many.retire_gracefully(pool); by_many = [np.array(r2.output(o).as_bytes(), copy=True) for o in (o1, o2)]; r2.finish()
    assert all(np.array_equal(x, y) for x, y in zip(by_one, by_many))

    ident = CommandBuffer()
    j = ident.input(bg.descriptor())
    oj, _ = ident.output(j)
    exe = Linker.from_included().compile(ident).lower_to(Capabilities.from_device(next(pool.iter_devices())))
    env = exe.from_pool(pool); env.bind(j, bg.key())
    ex = exe.launch(env)
    assert not ex.is_running() and ex.kernel_count() == 0
    ex.rerun().block_on()  # zos_program_launch left `running` set for an empty schedule before (advisor, round 1)
    ret = ex.retire_gracefully(pool)
    assert np.array_equal(ret.output(oj).as_bytes(), bg.as_bytes())
    ret.finish()


def test_batched_program_from_a_command_buffer(pool):
    """BASELINE config 4 is '256 frames': the same CommandBuffer lowered with Capabilities.batch = N runs N frames per
    kernel launch; every frame equals the single-frame run (frames are independent)."""
    rng = np.random.default_rng(5)
    W, H, N = 96, 64, 5
    desc = Descriptor.with_srgb_image("rgba8", W, H)
    frames_a = rng.integers(0, 256, (N, H * W * 4), dtype=np.uint8)
    frames_b = rng.integers(0, 256, (N, H * W * 4), dtype=np.uint8)
    c = CommandBuffer()
    below, above = c.input(desc), c.input(desc)
    out, _ = c.output(c.blend(below, Rectangle(0, 0, W, H), above, Blend.Alpha))
    plan = Linker.from_included().compile(c)
    ctx = next(pool.iter_devices())
    batched = plan.lower_to(Capabilities.from_device(ctx, batch=N))
    ka, kb = pool.insert(desc, frames_a, batch=N), pool.insert(desc, frames_b, batch=N)
    env = batched.from_pool(pool); env.bind(below, kb.key()); env.bind(above, ka.key())
    ex = batched.launch(env)
    assert ex.kernel_count() == 1
    while ex.is_running():
        ex.step().block_on()
    ret = ex.retire_gracefully(pool); got = np.array(ret.output(out).as_bytes(), copy=True).reshape(N, -1); ret.finish()
    od = O.srgb_rgba8(W, H)
    for f in range(N):
        exp = O.blend(O.Image(od, frames_b[f].reshape(H, W * 4)), (0, 0, W, H), O.Image(od, frames_a[f].reshape(H, W * 4)), 3)
        assert np.array_equal(got[f], exp.data.reshape(-1)), f
    single = plan.lower_to(Capabilities.from_device(ctx))
    env = single.from_pool(pool)
    with pytest.raises(StartError):  # a 5-frame pool image does not fit a single-frame program
        env.bind(below, kb.key())


def test_planar_output_is_sized_for_all_planes(pool):
    """An output that is a planar 4:2:0 register: the host image holds Y plus both chroma planes (1.5 bytes per pixel), not
    width * height * texel_stride (advisor, round 1: that was a heap overflow in Retire.output)."""
    W, H = 64, 48
    yuv = Z.yuv420_descriptor(W, H, Color.Rgb(Z.Primaries.Bt709, Transfer.Bt709), Z.YuvMatrix.Bt709, False, False, 0)
    rng = np.random.default_rng(3)
    data = rng.integers(16, 236, W * H * 3 // 2, dtype=np.uint8)
    src = pool.insert(yuv, data)
    c = CommandBuffer()
    i = c.input(yuv)
    o, _ = c.output(i)
    exe = Linker.from_included().compile(c).lower_to(Capabilities.from_device(next(pool.iter_devices())))
    env = exe.from_pool(pool); env.bind(i, src.key())
    ex = exe.launch(env)
    ret = ex.retire_gracefully(pool)
    out = ret.output(o)
    assert out.as_bytes().size == W * H * 3 // 2 and np.array_equal(out.as_bytes(), data)
    ret.finish()
