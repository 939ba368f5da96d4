"""GPU cases for the two parts of the function support that were written after the round's GPU budget was spent: a
generic ENTRY POINT run through Execution / Retire (register translation) and a function invoked from inside a template.
Their host side is covered on the CPU (tests/test_host_layer.py: the linked streams equal the flat pipeline's); these run
the same pipelines on the device and compare bytes with the flat pipeline.  (Seen passing on a B200 at the end of round 1:
GPUTEST_r01 lists it as XPASS; the xfail guard it carried until then is gone.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from zosimos_b200.buffer import Color, Descriptor, SampleParts, Texel  # noqa: E402
from zosimos_b200.command import CommandBuffer, InvocationArguments, Linker, Rectangle  # noqa: E402
from zosimos_b200.program import Capabilities, Pool  # noqa: E402


def srgb(w, h):
    return Descriptor.with_srgb_image("rgba8", w, h)


def run(plan, pool, binds, out_reg):
    executable = plan.lower_to(Capabilities.from_device(next(pool.iter_devices())))
    env = executable.from_pool(pool)
    for reg, key in binds:
        env.bind(reg, key)
    execution = executable.launch(env)
    while execution.is_running():
        execution.step().block_on()
    retire = execution.retire_gracefully(pool)
    img = retire.output(out_reg)
    retire.finish()
    return np.array(img.as_bytes(), copy=True)


def test_generic_entry_point_with_nested_call_equals_flat_pipeline():
    pool = Pool()
    pool.request_device(0)
    try:
        rng = np.random.default_rng(9)
        big = pool.insert(srgb(320, 200), rng.integers(0, 256, 320 * 200 * 4, dtype=np.uint8))
        small = pool.insert(srgb(64, 64), rng.integers(0, 256, 64 * 64 * 4, dtype=np.uint8))
        rgba8 = Texel.new_u8(SampleParts.RgbA)

        helper = CommandBuffer()                 # helper<T>(x: T) -> BT.709 transfer
        hv = helper.generic()
        helper.output(helper.color_convert(helper.input_generic(hv), Color.BT709_RGB, rgba8))
        main = CommandBuffer()                   # main<T>(image: T, small): helper<T>(image), then inscribe
        t = main.generic()
        f = main.function(helper.computed_signature())
        image, little = main.input_generic(t), main.input(srgb(64, 64))
        (conv,) = main.invoke(f, InvocationArguments(generics=[t], arguments=[image]))
        out, _ = main.output(main.inscribe(conv, Rectangle(0, 0, 64, 64), main.color_convert(little, Color.BT709_RGB, rgba8)))
        linker = Linker.from_included()
        got = run(linker.link(main, [srgb(320, 200)], [helper], [[1], []]), pool, [(image, big.key()), (little, small.key())], out)

        flat = CommandBuffer()
        i2, s2 = flat.input(srgb(320, 200)), flat.input(srgb(64, 64))
        c2 = flat.color_convert(i2, Color.BT709_RGB, rgba8)
        o2, _ = flat.output(flat.inscribe(c2, Rectangle(0, 0, 64, 64), flat.color_convert(s2, Color.BT709_RGB, rgba8)))
        exp = run(linker.compile(flat), pool, [(i2, big.key()), (s2, small.key())], o2)
        assert got.shape == exp.shape and np.array_equal(got, exp)
    finally:
        for c in pool.iter_devices():
            c.close()
